"""Kernel micro-benchmarks on cuda:0 (diagnostics, not the bench contract):
  gemm   : tvb_debug_gemm_bf16 (CTA-pair, cluster-of-4 multicast, single-CTA) vs torch.matmul on the configs[1] shapes,
           as a short burst (cool chip) and as a ~1.5 s sustained loop (power cap), CUDA-event timed.
  attn   : tvb_debug_attention v3/v4/v5 at the configs[1] shape + the clock64() trace of attention v4's first CTAs.
"""
import argparse
import ctypes as C
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from usdm_b200 import engine  # noqa: E402

lib = engine.load_library()
P = lambda t: C.c_void_p(t.data_ptr())
ST = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)


def timed(fn, iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def bench_gemm(out):
    M = 36864
    shapes = [("qkv", 3072, 1024, 0), ("attn_out", 1024, 1024, 2), ("ffn_in", 4096, 1024, 1), ("ffn_out", 1024, 4096, 2)]
    for name, N, K, epi in shapes:
        a = (torch.randn(M, K, device="cuda")).to(torch.bfloat16)
        w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
        bias = torch.randn(N, device="cuda")
        resid = torch.randn(M, N, device="cuda")
        of = torch.empty(M, N, device="cuda")
        ob = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        flops = 2.0 * M * N * K
        variants = {
            "pair": lambda: engine.check(lib.tvb_debug_gemm_bf16(0, P(a), P(w), P(bias), P(resid), P(of), P(ob), M, N, K, epi | 0x100, ST())),
            "pair_plain_bf16": lambda: engine.check(lib.tvb_debug_gemm_bf16(0, P(a), P(w), P(bias), P(resid), P(of), P(ob), M, N, K, 0 | 0x100, ST())),
            "mc4": lambda: engine.check(lib.tvb_debug_gemm_bf16(0, P(a), P(w), P(bias), P(resid), P(of), P(ob), M, N, K, epi | 0x200, ST())),
            "single": lambda: engine.check(lib.tvb_debug_gemm_bf16(0, P(a), P(w), P(bias), P(resid), P(of), P(ob), M, N, K, epi, ST())),
            "torch": lambda: torch.matmul(a, w.t()),
        }
        for vn, fn in variants.items():
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            import time
            time.sleep(0.3)                        # let the chip cool: burst number
            burst = timed(fn, 10)
            iters = max(20, int(1.5 / burst))
            sus = timed(fn, iters)
            out.append(dict(kind="gemm", shape=name, N=N, K=K, variant=vn, burst_tflops=flops / burst / 1e12, sustained_tflops=flops / sus / 1e12,
                            burst_us=burst * 1e6, sustained_us=sus * 1e6))
            print(out[-1], flush=True)


def bench_attn(out):
    H, NH = 1024, 16
    lens = [1134] * 32
    starts, o = [], 0
    for L in lens:
        starts.append(o)
        o += L + 16
    rows = (o + 511) // 512 * 512
    qkv = (torch.randn(rows, 3 * H, device="cuda") * 0.5).to(torch.bfloat16)
    ctx = torch.zeros(rows, H, device="cuda", dtype=torch.bfloat16)
    start = 2 ** (-2 ** -(math.log2(NH) - 3))
    slopes = torch.tensor([start * start ** i for i in range(NH)], dtype=torch.float32, device="cuda")
    s_arr = (C.c_int32 * len(lens))(*starts)
    l_arr = (C.c_int32 * len(lens))(*lens)
    flops = 4.0 * 64 * NH * sum(l * l for l in lens)

    def run(version, poly):
        engine.check(lib.tvb_debug_attention(0, P(qkv), P(ctx), s_arr, l_arr, len(lens), rows, H, NH, P(slopes), version, poly, ST()))

    for tv in (7,):
      # trace of v4 / v7 (the hook's own per-call overhead makes wall-clock timing of single launches meaningless here)
      for _ in range(3):
          run(tv, 0)
      trace = torch.zeros(2 * 16 * 256, dtype=torch.int64, device="cuda")
      engine.check(lib.tvb_debug_set_trace(P(trace)))
      run(tv, 0)
      torch.cuda.synchronize()
      engine.check(lib.tvb_debug_set_trace(None))
      trf = trace.cpu()[16 * 256:].view(16, 256)
      tr = trace.cpu()[:16 * 256].view(16, 256)
      for cta in (0, 5):
          row = tr[cta].tolist()
          t0 = row[0]
          rel = lambda v: (v - t0) if v else None
          d = dict(kind="attn_trace", version=tv, cta=cta, setup=rel(row[1]), q_landed=rel(row[176]),
                   A_s_ready=[rel(row[16 + 2 * j]) for j in range(9)], A_p_done=[rel(row[17 + 2 * j]) for j in range(9)], A_end=rel(row[80]),
                   B_s_ready=[rel(row[96 + 2 * j]) for j in range(9)], B_p_done=[rel(row[97 + 2 * j]) for j in range(9)], B_end=rel(row[160]),
                   mma_p_recv=[rel(row[177 + n]) for n in range(18)])
          if tv == 7:
              fr = trf[cta].tolist()
              d["A_fine"] = [[rel(fr[4 * j + k]) for k in range(4)] for j in range(9)]
              d["B_fine"] = [[rel(fr[128 + 4 * j + k]) for k in range(4)] for j in range(9)]
          out.append(d)
          print(d, flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="gemm,attn")
    ap.add_argument("--out", default="gpurun_out/microbench.json")
    a = ap.parse_args()
    res = []
    if "attn" in a.what:
        bench_attn(res)
    if "gemm" in a.what:
        bench_gemm(res)
    json.dump(res, open(a.out, "w"), indent=1)
