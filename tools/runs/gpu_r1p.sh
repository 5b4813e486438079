#!/bin/bash
# one gpurun call: attn8 with alternating PV barriers + next-block prefetch
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_attention.py -x -q -s -k "8-0 or 8-3 or long_form or persistent or peaky" > gpurun_out/t_attn8.log 2>&1; rc8=$?; echo "attn8 exit $rc8" > gpurun_out/summary.txt
if [ $rc8 -ne 0 ]; then cat gpurun_out/summary.txt; tail -20 gpurun_out/t_attn8.log; exit 0; fi
timeout 300 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_attention.py -q > gpurun_out/t_bf16.log 2>&1; echo "bf16+attn exit $?" >> gpurun_out/summary.txt
i=0
for cfg in "attn_poly=3" "attn_poly=0" "attn_poly=2" "attn_poly=4"; do
  timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --opt $cfg > "gpurun_out/bench_p$i.json" 2> "gpurun_out/bench_p$i.err"; echo "bench p$i [$cfg] exit $?" >> gpurun_out/summary.txt
  i=$((i+1))
done
cat gpurun_out/summary.txt; tail -3 gpurun_out/t_bf16.log
