"""Short workload for ncu: ONE BigVGAN forward at the configs[1] shape (16 x 861 frames) on cuda:0 (232 kernel launches)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from usdm_b200.vocoder import BigVGAN, BIGVGAN_22K_80  # noqa: E402
from tools.bench_vocoder import random_weights  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--frames", type=int, default=861)
ap.add_argument("--passes", type=int, default=1)
a = ap.parse_args()
m = BigVGAN(BIGVGAN_22K_80)
random_weights(m)
m = m.cuda().eval()
g = torch.Generator().manual_seed(1)
mel = (torch.randn(a.batch, 80, a.frames, generator=g) * 1.2).cuda()
for _ in range(a.passes):
    wav = m(mel)
torch.cuda.synchronize()
print("ok", float(wav.abs().mean()), "launches", m.kernel_launches())
