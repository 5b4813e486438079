#!/bin/bash
# one gpurun call: attention v8 (persistent + split converged issue warps)
mkdir -p gpurun_out
timeout 90 python -m pytest tests/test_gpu_attention.py -x -q -s -k "8-0 or 8-3 or long_form or persistent" > gpurun_out/t_attn8.log 2>&1; rc8=$?; echo "attn8 exit $rc8" > gpurun_out/summary.txt
if [ $rc8 -ne 0 ]; then cat gpurun_out/summary.txt; tail -20 gpurun_out/t_attn8.log; exit 0; fi
timeout 300 python -m pytest tests/test_gpu_attention.py -q > gpurun_out/t_attn.log 2>&1; echo "attn exit $?" >> gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_bf16.py -q -s > gpurun_out/t_bf16.log 2>&1; echo "bf16 exit $?" >> gpurun_out/summary.txt
i=0
for cfg in "attn_v=8" "attn_v=8 --opt attn_poly=0" "attn_v=8 --opt attn_poly=2" "attn_v=7"; do
  timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --opt $cfg > "gpurun_out/bench_j$i.json" 2> "gpurun_out/bench_j$i.err"; echo "bench j$i [$cfg] exit $?" >> gpurun_out/summary.txt
  i=$((i+1))
done
cat gpurun_out/summary.txt
