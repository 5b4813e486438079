#!/bin/bash
# one gpurun call: paced MUFU issue in attn8
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_attention.py -x -q -s -k "8-0 or 8-3 or 8-2 or 8-4 or long_form or persistent or peaky" > gpurun_out/t_attn8.log 2>&1; rc8=$?; echo "attn8 exit $rc8" > gpurun_out/summary.txt
timeout 300 python -m pytest tests/test_gpu_bf16.py -q > gpurun_out/t_bf16.log 2>&1; echo "bf16 exit $?" >> gpurun_out/summary.txt
i=0
for cfg in "attn_pace=1" "attn_pace=0" "attn_pace=1 --opt attn_poly=2" "attn_pace=1 --opt attn_poly=0" "attn_pace=1 --opt attn_poly=4"; do
  timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --opt $cfg > "gpurun_out/bench_q$i.json" 2> "gpurun_out/bench_q$i.err"; echo "bench q$i [$cfg] exit $?" >> gpurun_out/summary.txt
  i=$((i+1))
done
cat gpurun_out/summary.txt; tail -2 gpurun_out/t_attn8.log
