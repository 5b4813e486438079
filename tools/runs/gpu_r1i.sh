#!/bin/bash
# one gpurun call: elect_one role guards (uniform-datapath MMA issue) in GEMM/conv, defaults attn v7 poly 3, timing ablations of attn7
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_gemm.py -x -q > gpurun_out/t_gemm.log 2>&1; echo "gemm exit $?" > gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_fp32.py -q -s > gpurun_out/t_bf16.log 2>&1; echo "bf16+fp32 exit $?" >> gpurun_out/summary.txt
i=0
for cfg in "attn_poly=3" "attn_poly=101" "attn_poly=102" "attn_poly=104" "attn_poly=107"; do
  timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --opt $cfg > "gpurun_out/bench_i$i.json" 2> "gpurun_out/bench_i$i.err"; echo "bench i$i [$cfg] exit $?" >> gpurun_out/summary.txt
  i=$((i+1))
done
cat gpurun_out/summary.txt
