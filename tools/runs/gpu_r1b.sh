#!/bin/bash
# one gpurun call: kernel-level attention test, bf16 parity, then bench ablations
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_attention.py -x -q -s > gpurun_out/t_attn.log 2>&1; echo "attn exit $?" > gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_bf16.py -q -s > gpurun_out/t_bf16.log 2>&1; echo "bf16 exit $?" >> gpurun_out/summary.txt
for cfg in "default" "attn_poly=2" "attn_poly=3" "attn_poly=4" "attn_v=3" "ln_fuse=0"; do
  if [ "$cfg" = "default" ]; then opt=""; else opt="--opt $cfg"; fi
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline $opt > "gpurun_out/bench_${cfg}.json" 2> "gpurun_out/bench_${cfg}.err"; echo "bench $cfg exit $?" >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt
