"""Throughput of the vocoder path beside bench.py (which stays on the north-star mel-generator metric).

    python tools/bench_vocoder.py [--batch 16] [--seconds 10] [--steps 5] [--warmup 3] [--precision bf16] [--cpu-baseline]

Workload: the mels configs[1] hands to the vocoder (model_util.py:94,103-104): 16 utterances x 861 target frames, 80 bands,
synthetic N(0, 1.2) mels, seeded random weights of the published bigvgan_22khz_80band architecture (112 M parameters).
Prints one JSON line: waveform samples/s and mel frames/s through BigVGAN.forward (device-resident and host-to-host),
the algorithmic conv FLOP rate, and the per-category device time of one forward.
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from usdm_b200.vocoder import BigVGAN, BIGVGAN_22K_80  # noqa: E402
from bench import frames_for_seconds, load_peaks  # noqa: E402


def conv_flops_per_frame(h):
    """2 x MACs of every convolution per mel frame (conv_pre, upsamplers, AMP blocks, conv_post); activations not counted."""
    c0, rate, macs = h["upsample_initial_channel"], 1, h["num_mels"] * h["upsample_initial_channel"] * 7
    for i, (u, k) in enumerate(zip(h["upsample_rates"], h["upsample_kernel_sizes"])):
        cin, cout = c0 >> i, c0 >> (i + 1)
        macs += rate * cin * cout * k            # every input step meets all k taps of the transposed conv
        rate *= u
        macs += rate * cout * cout * sum(h["resblock_kernel_sizes"]) * 6
    macs += rate * (c0 >> len(h["upsample_rates"])) * 7
    return 2 * macs


def random_weights(model, seed=0):
    g = torch.Generator().manual_seed(seed)
    sd = model.state_dict()
    for k, v in sd.items():
        if k.endswith(".weight"):
            fan = v[0].numel() if "ups." not in k else v.shape[0] * v.shape[2] / 4
            sd[k] = torch.randn(v.shape, generator=g) / fan ** 0.5 * (0.2 if k.startswith("conv_post") else (0.5 if "convs2" in k else 1.0))
        elif k.endswith(".bias"):
            sd[k] = 0.1 * torch.randn(v.shape, generator=g)
        elif k.endswith((".alpha", ".beta")):
            sd[k] = 0.5 * torch.randn(v.shape, generator=g)
    model.load_state_dict(sd)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--cpu-baseline", action="store_true")
    ap.add_argument("--opt", action="append", default=[], help="vocoder option name=value (e.g. halo=0)")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    h = BIGVGAN_22K_80
    model = BigVGAN(h)
    random_weights(model)
    model = model.to(dev).eval().set_precision(args.precision)
    for kv in args.opt:
        k, v = kv.split("=")
        model.set_option(k, int(v))
    B, T = args.batch, frames_for_seconds(args.seconds)
    g = torch.Generator().manual_seed(1)
    mel_host = (torch.randn(B, 80, T, generator=g) * 1.2).pin_memory()
    mel = mel_host.to(dev)
    wav_host = torch.empty(B, 1, T * 256).pin_memory()
    for _ in range(max(args.warmup, 3)):
        model(mel)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = model.kernel_launches()
    e0.record()
    for _ in range(args.steps):
        wav = model(mel)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    launches = (model.kernel_launches() - l0) // args.steps
    t0 = time.perf_counter()
    for _ in range(args.steps):
        wav_host.copy_(model(mel_host.to(dev, non_blocking=True)), non_blocking=True)
        torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) / args.steps * 1e3
    model.set_profile(True)
    model(mel)
    prof = model.read_profile()
    model.set_profile(False)
    fl = conv_flops_per_frame(h) * B * T
    peaks = load_peaks()
    line = dict(metric="vocoder_wav_samples_per_sec", value=B * T * 256 / (ms / 1e3), unit="samples/s", mel_frames_per_s=B * T / (ms / 1e3),
                ms_per_forward=ms, options=args.opt, n_gpus=1, steps=args.steps, dtype=args.precision, data="synthetic",
                config=dict(workload=f"BigVGAN 22 kHz/80-band forward, {B} x {T} frames ({args.seconds:g} s each)", params_m=sum(p.numel() for p in model.parameters()) / 1e6),
                e2e=dict(value=B * T * 256 / (e2e_ms / 1e3), unit="samples/s", h2d_bytes_per_step=mel_host.numel() * 4, d2h_bytes_per_step=wav_host.numel() * 4),
                gpu_launches=int(launches), conv_tflops=fl / (ms / 1e3) / 1e12, kernels_ms=prof,
                roofline=dict(bound="tensor", kernel="voc_conv_tc_kernel", achieved=fl / (prof["conv"] / 1e3) / 1e12, peak=peaks["sustained"], unit="TFLOP/s",
                              frac=fl / (prof["conv"] / 1e3) / 1e12 / peaks["sustained"]),
                device_gb=model.device_bytes() / 2 ** 30, realtime_factor=B * T * 256 / 22050 / (ms / 1e3))
    if args.cpu_baseline:
        from oracle import vocoder_oracle as V
        sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        melc = mel_host[:1, :, :64].clone()
        V.bigvgan_forward(sd, melc[:, :, :8], h)
        t0 = time.perf_counter()
        V.bigvgan_forward(sd, melc, h)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = dict(value=64 * 256 / dt, unit="samples/s", cores=torch.get_num_threads(), kind="port", sample="1 utterance x 64 frames, torch fp32 oracle")
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
