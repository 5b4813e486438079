"""Single-request latency through the public API: glue.reconstruct_speech (model_util.py:72-105) with the reference app's settings
(n_timesteps=50, Heun -> 49 estimator evaluations with CFG, speech prompt from a 3.16 s reference clip), one utterance.

    python tools/latency.py [--seconds 5] [--reps 5]

Synthetic units / reference audio, random-init weights of the real architectures.  Prints one JSON line (wall clock, host to host).
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import usdm_b200  # noqa: E402
from usdm_b200 import glue  # noqa: E402
from bench import MODEL_CFG  # noqa: E402
from tools.bench_vocoder import random_weights  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=5.0)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--n-timesteps", type=int, default=50)
    a = ap.parse_args()
    torch.manual_seed(0)
    vb = usdm_b200.Voicebox(**MODEL_CFG).cuda().eval()
    voc = usdm_b200.BigVGAN()
    random_weights(voc)
    voc = voc.cuda().eval()
    g = torch.Generator().manual_seed(1)
    agent = torch.randint(0, 10000, (int(a.seconds * 50),), generator=g)
    ref_units = torch.randint(0, 10000, (158,), generator=g)                 # 3.16 s of 50 Hz units -> 272 prompt frames
    ref_audio = (0.1 * torch.randn(int(3.16 * 22050), generator=g)).clamp(-1, 1)
    times = []
    for i in range(a.reps + 2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        wav = glue.reconstruct_speech(agent, vb, voc, ref_audio, ref_units, n_timesteps=a.n_timesteps)   # returns a host numpy array
        dt = time.perf_counter() - t0
        if i >= 2:
            times.append(dt)
    times.sort()
    med = times[len(times) // 2]
    print(json.dumps(dict(metric="reconstruct_speech latency", value=med * 1e3, unit="ms", utterance_seconds=len(wav) / 22050, n_timesteps=a.n_timesteps,
                          solver="heun", nfe=2 * ((a.n_timesteps + 1) // 2) - 1, prompt_frames=272, real_time_factor=len(wav) / 22050 / med,
                          all_ms=[round(t * 1e3, 1) for t in times], data="synthetic", dtype="bf16")), flush=True)


if __name__ == "__main__":
    main()
