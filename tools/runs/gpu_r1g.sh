#!/bin/bash
# one gpurun call: persistent attention v6 (after the half-B phase fix) + 3-buffer GEMM epilogue
mkdir -p gpurun_out
timeout 90 python -m pytest tests/test_gpu_attention.py -x -q -s -k "6-0 or 6-3 or v6 or long_form" > gpurun_out/t_attn6.log 2>&1; rc6=$?; echo "attn6 exit $rc6" > gpurun_out/summary.txt
timeout 200 python -m pytest tests/test_gpu_gemm.py -x -q > gpurun_out/t_gemm.log 2>&1; echo "gemm exit $?" >> gpurun_out/summary.txt
v=4; if [ $rc6 -eq 0 ]; then v=6; fi
timeout 300 python -m pytest tests/test_gpu_attention.py -q -s > gpurun_out/t_attn.log 2>&1; echo "attn exit $?" >> gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_bf16.py -q -s > gpurun_out/t_bf16.log 2>&1; echo "bf16 exit $?" >> gpurun_out/summary.txt
i=0
for cfg in "attn_v=$v" "attn_v=$v --opt attn_poly=3" "attn_v=$v --opt attn_poly=4" "attn_v=4 --opt attn_poly=3"; do
  timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --opt $cfg > "gpurun_out/bench_g$i.json" 2> "gpurun_out/bench_g$i.err"; echo "bench g$i [$cfg] exit $?" >> gpurun_out/summary.txt
  i=$((i+1))
done
cat gpurun_out/summary.txt
