"""Throughput of BASELINE.json configs[0..4] on the local GPU(s) (diagnostics beside bench.py, which stays on configs[1]).

    python tools/sweep.py                       # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/sweep.py

Synthetic units / prompts / noise, random-init weights, 32-step Euler + CFG (g = 1), 272-frame prompt everywhere.
config 4 / 5 lengths: LogNormal(ln 4.5 s, 0.6) clipped to [2, 20] s, numpy default_rng(1234); config 5 is sharded
over the ranks by usdm_b200.parallel (LPT on the FLOP model, micro-batches of <= 80 k CFG tokens, one all-gather of the mels).
Prints one JSON line per config: target frames/s (whole job), seconds, and the valid-token FLOP rate per GPU.
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import usdm_b200  # noqa: E402
from usdm_b200 import flops as F, parallel  # noqa: E402
from bench import MODEL_CFG, frames_for_seconds, load_peaks  # noqa: E402

P = 272
NFE = 32


def batch(lens_target, seed, dev, P=P):
    g = torch.Generator().manual_seed(seed)
    lens = [t + P for t in lens_target]
    B, T = len(lens), max(lens)
    x = torch.randint(0, 10000, (B, T), generator=g)
    cond = torch.zeros(B, 80, T)
    cond[:, :, :P] = torch.randn(B, 80, P, generator=g)
    return x.to(dev), cond.to(dev), torch.tensor(lens)


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    model = usdm_b200.Voicebox(**MODEL_CFG).eval()
    for i, a in enumerate(sys.argv):                   # --opt name=value engine options (ablations), as bench.py
        if a == "--opt" and i + 1 < len(sys.argv):
            k, v = sys.argv[i + 1].split("=")
            model.set_option(k, int(v))
    peaks = load_peaks()
    plen = torch.tensor([P])

    def gen(x, cond, L):
        return model.generate(x, cond, L, NFE, solver="euler", gradient_scale=1.0, speech_prompt=True, prompt_lengths=plen)

    def gen_noprompt(x, cond, L):
        return model.generate(x, cond, L, NFE, solver="euler", gradient_scale=1.0, speech_prompt=False)

    def timed(fn, reps):
        fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        return (time.perf_counter() - t0) / reps

    rng = np.random.default_rng(1234)
    d64 = np.clip(rng.lognormal(np.log(4.5), 0.6, 64), 2.0, 20.0)
    rng = np.random.default_rng(1234)
    d512 = np.clip(rng.lognormal(np.log(4.5), 0.6, 512), 2.0, 20.0)
    cases = [
        ("configs[0]-shape: 1 utterance x 5 s (latency)", [frames_for_seconds(5.0)], 5, False),
        ("configs[1]: 16 x 10 s", [frames_for_seconds(10.0)] * 16, 3, False),
        ("configs[2]: 8 x 30 s long-form", [frames_for_seconds(30.0)] * 8, 2, False),
        ("configs[3]: ragged 64, 2-20 s", [frames_for_seconds(float(s)) for s in d64], 2, False),
        ("configs[4]: 512 ragged utterances, sharded", [frames_for_seconds(float(s)) for s in d512], 1, True),
        ("configs[1] without the prompt: 16 x 10 s, S = 862 (SURVEY 8d asks for both)", [frames_for_seconds(10.0)] * 16, 3, False),
    ]
    only = os.environ.get("SWEEP_ONLY")             # substring filter, e.g. SWEEP_ONLY="without the prompt"
    for name, lens_t, reps, sharded in cases:
        if only and only not in name:
            continue
        noprompt = "without the prompt" in name
        Pc = 0 if noprompt else P
        x, cond, L = batch(lens_t, 7, dev, Pc)
        if sharded:
            fn = lambda: parallel.generate_sharded(gen, x, cond, L, max_tokens=80000)
            tokens = [t + Pc + 1 for t in lens_t] * 2
            frames = sum(lens_t)
            per_gpu_flops = F.estimator_flops(tokens, MODEL_CFG)["total"] * NFE / world
        else:
            fn = (lambda: gen_noprompt(x, cond, L)) if noprompt else (lambda: gen(x, cond, L))   # every rank runs the same batch (replicas)
            tokens = [t + Pc + 1 for t in lens_t] * 2
            frames = sum(lens_t) * world
            per_gpu_flops = F.estimator_flops(tokens, MODEL_CFG)["total"] * NFE
        sec = timed(fn, reps)
        if rank == 0:
            print(json.dumps(dict(config=name, n_gpus=world, utterances=len(lens_t), target_frames=frames, seconds=sec,
                                  frames_per_s=frames / sec, tflops_per_gpu=per_gpu_flops / sec / 1e12,
                                  frac_of_sustained_bf16=per_gpu_flops / sec / 1e12 / peaks["sustained"])), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
