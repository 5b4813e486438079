"""Short workload for ncu: a few CFG estimator passes at BASELINE configs[1] size (B=16, T=861+272) on cuda:0."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import usdm_b200  # noqa: E402
from bench import MODEL_CFG  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--frames", type=int, default=1133)
ap.add_argument("--passes", type=int, default=2)
ap.add_argument("--layers", type=int, default=24)
ap.add_argument("--opt", action="append", default=[])
a = ap.parse_args()
cfg = dict(MODEL_CFG, num_hidden_layers=a.layers)
torch.manual_seed(0)
m = usdm_b200.Voicebox(**cfg).eval()
for kv in a.opt:
    k, v = kv.split("=")
    m.set_option(k, int(v))
g = torch.Generator().manual_seed(0)
B, T = a.batch, a.frames
x = torch.randint(0, 10000, (B, T), generator=g).cuda()
cond = torch.randn(B, 80, T, generator=g).cuda()
z = torch.randn(B, 80, T, generator=g).cuda()
L = torch.full((B,), T)
for i in range(a.passes):
    v, _ = m.sample(x, z, cond, L, torch.tensor(0.3), 1.0, True)
torch.cuda.synchronize()
print("ok", float(v.abs().mean()), "launches", m.kernel_launches())
