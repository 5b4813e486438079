#!/bin/bash
# one gpurun call: persistent attention v6 + glue kernels + full-size config tests + bench
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_attention.py -x -q -s -k "6-0 or 6-3 or v6 or long_form" > gpurun_out/t_attn6.log 2>&1; rc6=$?; echo "attn6 exit $rc6" > gpurun_out/summary.txt
timeout 300 python -m pytest tests/test_gpu_glue.py -q -s > gpurun_out/t_glue.log 2>&1; echo "glue exit $?" >> gpurun_out/summary.txt
timeout 400 python -m pytest tests/test_gpu_configs.py -q -s > gpurun_out/t_configs.log 2>&1; echo "configs exit $?" >> gpurun_out/summary.txt
if [ $rc6 -eq 0 ]; then
  timeout 300 python -m pytest tests/test_gpu_attention.py -q -s > gpurun_out/t_attn.log 2>&1; echo "attn exit $?" >> gpurun_out/summary.txt
  timeout 600 python -m pytest tests/test_gpu_bf16.py -q -s > gpurun_out/t_bf16.log 2>&1; echo "bf16 exit $?" >> gpurun_out/summary.txt
  i=0
  for cfg in "attn_v=6 --opt gemm_mc=0" "attn_v=6 --opt attn_poly=3 --opt gemm_mc=0" "attn_v=6 --opt attn_poly=4 --opt gemm_mc=0" "attn_v=6 --opt attn_poly=2 --opt gemm_mc=0"; do
    timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --opt $cfg > "gpurun_out/bench_f$i.json" 2> "gpurun_out/bench_f$i.err"; echo "bench f$i [$cfg] exit $?" >> gpurun_out/summary.txt
    i=$((i+1))
  done
fi
cat gpurun_out/summary.txt
