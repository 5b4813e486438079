#!/bin/bash
# one gpurun call: attention v7 (two issue warps): tests, trace, bench vs v4 on the same box
mkdir -p gpurun_out
timeout 90 python -m pytest tests/test_gpu_attention.py -x -q -s -k "7-0 or 7-3 or long_form" > gpurun_out/t_attn7.log 2>&1; rc7=$?; echo "attn7 exit $rc7" > gpurun_out/summary.txt
if [ $rc7 -ne 0 ]; then cat gpurun_out/summary.txt; tail -20 gpurun_out/t_attn7.log; exit 0; fi
timeout 120 python tools/microbench.py --what attn > gpurun_out/microbench.log 2>&1; echo "micro exit $?" >> gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_bf16.py -q -s -k "attn_v" > gpurun_out/t_bf16.log 2>&1; echo "bf16 exit $?" >> gpurun_out/summary.txt
i=0
for cfg in "attn_v=7" "attn_v=7 --opt attn_poly=3" "attn_v=7 --opt attn_poly=2" "attn_v=4 --opt attn_poly=3"; do
  timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --opt $cfg > "gpurun_out/bench_h$i.json" 2> "gpurun_out/bench_h$i.err"; echo "bench h$i [$cfg] exit $?" >> gpurun_out/summary.txt
  i=$((i+1))
done
cat gpurun_out/summary.txt
