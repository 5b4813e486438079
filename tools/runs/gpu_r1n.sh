#!/bin/bash
# one gpurun call: zigzag row-block traversal (L2 reuse between producer and consumer kernels): parity + A/B bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_gemm.py -x -q > gpurun_out/t_zz.log 2>&1; echo "parity exit $?" > gpurun_out/summary.txt
i=0
for cfg in "zigzag=1" "zigzag=0" "zigzag=1" "zigzag=0"; do
  timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --opt $cfg > "gpurun_out/bench_n$i.json" 2> "gpurun_out/bench_n$i.err"; echo "bench n$i [$cfg] exit $?" >> gpurun_out/summary.txt
  i=$((i+1))
done
cat gpurun_out/summary.txt; tail -2 gpurun_out/t_zz.log
