#!/bin/bash
# one gpurun call: multicast GEMM tests + attention v5 (two threads per row) tests, parity, bench ablations
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm.py -x -q > gpurun_out/t_gemm.log 2>&1; echo "gemm exit $?" > gpurun_out/summary.txt
timeout 200 python -m pytest tests/test_gpu_attention.py -x -q -s -k "5-" > gpurun_out/t_attn5.log 2>&1; rc5=$?; echo "attn5 exit $rc5" >> gpurun_out/summary.txt
timeout 300 python -m pytest tests/test_gpu_attention.py -q -s > gpurun_out/t_attn.log 2>&1; echo "attn exit $?" >> gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_bf16.py -q -s > gpurun_out/t_bf16.log 2>&1; echo "bf16 exit $?" >> gpurun_out/summary.txt
cfgs=("attn_v=4 --opt attn_poly=3" "attn_v=4 --opt attn_poly=3 --opt gemm_mc=0")
if [ $rc5 -eq 0 ]; then cfgs+=("attn_v=5" "attn_v=5 --opt attn_poly=4" "attn_v=5 --opt attn_poly=3" "attn_v=5 --opt attn_poly=2"); fi
i=0
for cfg in "${cfgs[@]}"; do
  timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --opt $cfg > "gpurun_out/bench_e$i.json" 2> "gpurun_out/bench_e$i.err"; echo "bench e$i [$cfg] exit $?" >> gpurun_out/summary.txt
  i=$((i+1))
done
cat gpurun_out/summary.txt
