#!/bin/bash
# one gpurun call: attention v4 (pipelined softmax, per-buffer barriers): kernel test, parity, bench ablations, ncu of the attention kernel
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_attention.py -x -q -s > gpurun_out/t_attn.log 2>&1; rc=$?; echo "attn exit $rc" > gpurun_out/summary.txt
if [ $rc -ne 0 ]; then cat gpurun_out/summary.txt; exit 0; fi
timeout 600 python -m pytest tests/test_gpu_bf16.py -q -s > gpurun_out/t_bf16.log 2>&1; echo "bf16 exit $?" >> gpurun_out/summary.txt
for cfg in "default" "attn_poly=4" "attn_poly=3" "attn_poly=2"; do
  if [ "$cfg" = "default" ]; then opt=""; else opt="--opt $cfg"; fi
  timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline $opt > "gpurun_out/bench_${cfg}.json" 2> "gpurun_out/bench_${cfg}.err"; echo "bench $cfg exit $?" >> gpurun_out/summary.txt
done
timeout 120 python tools/profile_pass.py --layers 2 --passes 2 > gpurun_out/plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn4 -s 2 -c 2 -o gpurun_out/prof_attn4 -f python tools/profile_pass.py --layers 2 --passes 2 > gpurun_out/ncu_attn4.log 2>&1
echo "ncu exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
