#!/usr/bin/env python
"""bench.py — token-Voicebox mel frames/s @ 32 NFE on B200 (BASELINE.json metric), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one complete solve of one batch: BASELINE configs[1] = 16 utterances x 10 s of units (861 target mel frames each)
behind a 272-frame reference-mel prompt, 32 Euler steps with classifier-free guidance (64 transformer passes over
2 x 16 sequences of 1134 tokens), bf16 tensor-core path, synthetic units / prompt / noise, random-init weights.
`value` = generated TARGET frames per second over all ranks with the inputs resident in HBM; `e2e` = the same through
Voicebox.generate() with HOST inputs (pinned) copied in and the mel copied out inside the timed region.
Multi-GPU: utterance batches shard over ranks (weak scaling, per-GPU batch fixed), no data-path collective; one NCCL
all-gather of the output mels per step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

MODEL_CFG = dict(n_feats=80, n_tokens=10000, embedding_dim=1280, hidden_size=1024, intermediate_size=4096,
                 num_attention_heads=16, num_hidden_layers=24, convpos_width=31, convpos_groups=16, convpos_depth=2,
                 attention_dropout=0.0, activation_dropout=0.1, hidden_dropout=0.0, solver="euler", sigma_min=1e-4)
METRIC = "token-Voicebox mel frames/s @32 NFE"
UNIT = "mel frames/s"


def frames_for_seconds(sec: float) -> int:
    return int(22050 * sec) // 256            # hop 256 @ 22.05 kHz (config.json:8-11)


def workload(args):
    T_target = frames_for_seconds(args.seconds)
    P = args.prompt_frames
    return dict(B=args.batch, T_target=T_target, P=P, T=T_target + P, nfe=args.nfe, gscale=1.0)


def make_inputs(w, seed, device):
    g = torch.Generator().manual_seed(seed)
    x = torch.randint(0, 10000, (w["B"], w["T"]), generator=g)
    cond = torch.zeros(w["B"], 80, w["T"])
    if w["P"] > 0:
        cond[:, :, :w["P"]] = torch.randn(w["B"], 80, w["P"], generator=g)
    lengths = torch.full((w["B"],), w["T"], dtype=torch.int64)
    return x.to(device), cond.to(device), lengths


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(sustained=float(d["bf16_tflops_sustained"]), burst=float(d["bf16_tflops"]), hbm=float(d["hbm_gbs"]), source="measured")
    return dict(sustained=1400.0, burst=1590.0, hbm=6650.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [c.strip() for c in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None,
                    power_w_max=max(pw) if pw else None, samples=len(sm), reasons=sorted(reasons))


# ------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port of the reference decoder, fp32, on the box's host cores
# ------------------------------------------------------------------------------------------------------------
def cpu_reference_sample(w, steps, warmup):
    """Times CFG estimator evaluations (= 1/nfe of a solve) for ONE utterance of the workload on all host threads.
    Returns (frames_per_s extrapolated to the full solve, seconds per evaluation, cores, description)."""
    from oracle import voicebox_oracle as O      # the one place bench.py executes oracle/: as the CPU baseline
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = dict(O.DEFAULT_MODEL_CFG)
    sd = O.init_reference_weights(0, cfg)
    g = torch.Generator().manual_seed(0)
    T = w["T"]
    x = torch.randint(0, 10000, (1, T), generator=g)
    cond = torch.zeros(1, 80, T)
    cond[:, :, :w["P"]] = torch.randn(1, 80, w["P"], generator=g)
    z = torch.randn(1, 80, T, generator=g)
    L = torch.tensor([T])
    ts = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            O.sample_cfg(sd, cfg, x, z, cond, L, torch.tensor(0.5), w["gscale"], True)
            if i >= warmup:
                ts.append(time.perf_counter() - t0)
    per_eval = sum(ts) / len(ts)
    fps = w["T_target"] / (per_eval * w["nfe"])
    sample = (f"{steps} CFG estimator evaluations (+{warmup} warm-up) of 1 utterance of the workload (T={T} frames incl. {w['P']}-frame "
              f"prompt, fp32 torch CPU oracle port); frames/s = {w['T_target']} target frames / ({w['nfe']} x mean evaluation time)")
    return fps, per_eval, cores, sample


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # One reference-arm "step" is the bounded sample itself: ONE CFG estimator evaluation of one utterance of the workload (1 / nfe of
    # that utterance's solve).  ms_per_step is its measured time (so steps x ms_per_step is the real timed region); `value` scales
    # it to the metric: target frames of the utterance / (nfe evaluations), which is batch-size independent on the CPU.
    steps = max(1, args.steps)
    warmup = max(0, args.warmup)
    fps, per_eval, cores, sample = cpu_reference_sample(w, steps, warmup)
    line = dict(metric=METRIC, value=fps, unit=UNIT, n_gpus=args.gpus, steps=steps, warmup=warmup,
                ms_per_step=per_eval * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic", impl="reference", config=bench_config(args, w),
                cpu_baseline=dict(value=fps, unit=UNIT, cores=cores, kind="port", sample=sample),
                e2e=dict(value=fps, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


def bench_config(args, w):
    return dict(workload=(f"BASELINE configs[1]: batch {w['B']} x {args.seconds:g} s unit sequences ({w['T_target']} target frames) + "
                          f"{w['P']}-frame mel prompt, {w['nfe']}-step Euler + CFG (g=1), per GPU"),
                batch_per_gpu=w["B"], target_frames=w["T_target"], prompt_frames=w["P"], tokens_per_pass=2 * w["B"] * (w["T"] + 1),
                nfe=w["nfe"], solver="euler", cfg_scale=w["gscale"], parallelism=f"dp{args.gpus} (utterance shards, all-gather of mels)",
                l2="not flushed: each pass streams >1 GB of activations, far above the 126 MB L2")


# ------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--prompt-frames", type=int, default=272)
    ap.add_argument("--nfe", type=int, default=32)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--no-vocoder", action="store_true", help="skip the extra (untimed-region) vocoder measurement reported under `vocoder`")
    ap.add_argument("--opt", action="append", default=[], help="engine option name=value (ablation), e.g. gemm_2cta=0")
    args = ap.parse_args()
    w = workload(args)

    if args.impl == "reference":
        run_reference(args, w)
        return

    import torch.distributed as dist
    import usdm_b200
    from usdm_b200 import flops as F

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (usdm_b200 has no CPU path); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if world != args.gpus and rank == 0:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE {world}", file=sys.stderr)

    torch.manual_seed(0)
    model = usdm_b200.Voicebox(**MODEL_CFG).eval()        # reference-identical seeded random init (tests/test_host_mirror.py)
    model.set_precision("bf16")
    for kv in args.opt:
        k, v = kv.split("=")
        model.set_option(k, int(v))
    x, cond, lengths = make_inputs(w, 1234 + rank, dev)
    plen = torch.tensor([w["P"]])
    sp = w["P"] > 0
    gathered = torch.empty((world, w["B"], 80, w["T"]), device=dev) if world > 1 else None

    def step_resident():
        out = model.generate(x, cond, lengths, w["nfe"], solver="euler", gradient_scale=w["gscale"], speech_prompt=sp, prompt_lengths=plen)
        if world > 1:
            dist.all_gather_into_tensor(gathered, out)
        return out

    x_host = x.cpu().pin_memory()
    cond_host = cond.cpu().pin_memory()
    out_host = torch.empty((w["B"], 80, w["T"]), dtype=torch.float32).pin_memory()

    def step_e2e():
        xd = x_host.to(dev, non_blocking=True)
        cd = cond_host.to(dev, non_blocking=True)
        out = model.generate(xd, cd, lengths, w["nfe"], solver="euler", gradient_scale=w["gscale"], speech_prompt=sp, prompt_lengths=plen)
        if world > 1:
            dist.all_gather_into_tensor(gathered, out)
        out_host.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return out_host

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms / 1e3, wall], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1])

    for _ in range(max(args.warmup, 3)):
        step_resident()
    l0 = model.kernel_launches()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    dev_s, _ = timed(step_resident, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    launches = model.kernel_launches() - l0
    step_e2e()
    _, e2e_wall = timed(step_e2e, args.steps)

    frames_per_step = w["B"] * w["T_target"] * world
    value = frames_per_step * args.steps / dev_s
    e2e_value = frames_per_step * args.steps / e2e_wall

    # ---- per-kernel split of one solve (CUDA events inside the library, on the launching stream) -> roofline
    peaks = load_peaks()
    fl = F.estimator_flops([w["T"] + 1] * (2 * w["B"]), MODEL_CFG)
    passes = w["nfe"]
    roof = None
    roof_hbm = None
    roof_attn = None
    kernels = None
    if not args.no_profile:
        model.set_option("profile", 1)
        step_resident()
        prof = model.read_profile()
        model.set_option("profile", 0)
        kernels = {k: dict(ms_per_step=v[0], launches=v[1]) for k, v in prof.items()}
        gemm_ms, gemm_n = prof["gemm"]
        achieved = fl["gemm"] * passes / (gemm_ms / 1e3) / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "gemm_traffic.json")       # DRAM bytes per launch from the committed ncu --set full capture
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get("avg_bytes_per_launch")
        roof = dict(bound="tensor", kernel="gemm2_tc_kernel (tcgen05 bf16 linear layers, 108 launches per pass)", achieved=achieved, peak=peaks["sustained"],
                    unit="TFLOP/s", frac=achieved / peaks["sustained"], traffic=traffic, peak_source=peaks["source"] + " bf16_tflops_sustained",
                    flops_per_launch=fl["gemm"] * passes / max(gemm_n, 1), avg_launch_ms=gemm_ms / max(gemm_n, 1),
                    share_of_step=gemm_ms / sum(v[0] for k, v in prof.items() if not k.startswith("gemm_")))
        for k in ("attention", "posconv"):
            ms = prof[k][0]
            if ms > 0:
                kernels[k]["tflops"] = fl[k] * passes / (ms / 1e3) / 1e12
        # the attention kernel against the same tensor peak (its algorithmic FLOPs: 4 d S^2 per head, skipped key blocks still count); it is
        # bound by the softmax side (MUFU.EX2 at 16 / clk / SM and the per-block hand-over chain), not by the tensor pipe
        at_ms, at_n = prof.get("attention", (0.0, 0))
        if at_ms > 0 and at_n > 0:
            # attention launches per pass (the planning kernels that share the category are counted in at_n): one per layer
            at_launch = MODEL_CFG["num_hidden_layers"] * passes
            at_tf = fl["attention"] * passes / (at_ms / 1e3) / 1e12
            roof_attn = dict(bound="tensor", kernel="attn10_kernel (fused QK^T + ALiBi + key mask + softmax + PV on tcgen05, 24 launches per pass)",
                             achieved=at_tf, peak=peaks["sustained"], unit="TFLOP/s", frac=at_tf / peaks["sustained"],
                             flops_per_launch=fl["attention"] * passes / at_launch, avg_launch_ms=at_ms / at_launch,
                             peak_source=peaks["source"] + " bf16_tflops_sustained",
                             note="softmax-bound (MUFU / hand-over chain): ncu 30 % tensor-pipe active, 46 % XU (profiles/r02/ncu_attn10_final_summary.txt)")
        # the row kernel of the path against the HBM roofline: LayerNorm reads the fp32 rows and writes the bf16
        # operand copy (6 B per element of the VALID tokens; per-row statistics and the dirty lines of the producer not counted)
        ln_ms, ln_n = prof.get("layernorm", (0.0, 0))
        if ln_ms > 0 and ln_n > 0:
            ln_bytes = 2 * w["B"] * (w["T"] + 1) * MODEL_CFG["hidden_size"] * 6
            gbs = ln_bytes / (ln_ms / ln_n / 1e3) / 1e9
            roof_hbm = dict(bound="hbm", kernel="k_layernorm (fp32 rows -> bf16 + row statistics, 25 launches per pass)", achieved=gbs, peak=peaks["hbm"],
                            unit="GB/s", frac=gbs / peaks["hbm"], bytes_per_launch=ln_bytes, avg_launch_ms=ln_ms / ln_n,
                            peak_source=peaks["source"] + " hbm_gbs")
    step_tflops = fl["total"] * passes * args.steps / dev_s / 1e12      # per GPU: fl counts this rank's batch, dev_s is the max over ranks

    cpu = None
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        fps, per_eval, cores, sample = cpu_reference_sample(w, 2, 1)
        cpu = dict(value=fps, unit=UNIT, cores=cores, kind="port", sample=sample)

    # ---- SURVEY.md §8 row f-3, reported beside the headline (never inside its timed region): the BigVGAN forward on the mels this
    # workload hands to the vocoder (model_util.py:94,104), and the resulting units -> waveform time of the whole batch
    voc = None
    if rank == 0 and world == 1 and not args.no_vocoder:
        from usdm_b200.vocoder import BigVGAN, BIGVGAN_22K_80
        from tools.bench_vocoder import conv_flops_per_frame, random_weights
        vmodel = BigVGAN(BIGVGAN_22K_80)
        random_weights(vmodel)
        vmodel = vmodel.to(dev).eval()
        vmel = torch.randn(w["B"], 80, w["T_target"], device=dev) * 1.2
        for _ in range(3):
            vmodel(vmel)
        torch.cuda.synchronize()
        v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = vmodel.kernel_launches()
        v0.record()
        for _ in range(3):
            vmodel(vmel)
        v1.record()
        torch.cuda.synchronize()
        vms = v0.elapsed_time(v1) / 3
        solve_ms = dev_s / args.steps * 1e3
        voc = dict(workload=f"BigVGAN 22 kHz/80-band forward on the {w['B']} x {w['T_target']} target frames of this workload (synthetic mel, random-init weights)",
                   ms_per_forward=vms, wav_samples_per_s=w["B"] * w["T_target"] * 256 / (vms / 1e3), gpu_launches=int((vmodel.kernel_launches() - l0) // 3),
                   conv_tflops=conv_flops_per_frame(BIGVGAN_22K_80) * w["B"] * w["T_target"] / (vms / 1e3) / 1e12,
                   units_to_waveform_ms=solve_ms + vms, units_to_waveform_frames_per_s=w["B"] * w["T_target"] / ((solve_ms + vms) / 1e3))
        del vmodel

    if rank == 0:
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
                    ms_per_step=dev_s / args.steps * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16",
                    data="synthetic", config=bench_config(args, w), clocks=clocks,
                    e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=int(x_host.numel() * 8 + cond_host.numel() * 4),
                             d2h_bytes_per_step=int(out_host.numel() * 4), timing="wall clock, max over ranks"),
                    gpu_launches=int(launches), roofline=roof, roofline_hbm=roof_hbm, roofline_attention=roof_attn,
                    roofline_step=dict(bound="tensor", achieved=step_tflops, peak=peaks["sustained"], unit="TFLOP/s per GPU",
                                       frac=step_tflops / peaks["sustained"],
                                       note="algorithmic FLOPs of the whole solve (valid tokens) / device time"),
                    kernels=kernels, cpu_baseline=cpu, vocoder=voc)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
