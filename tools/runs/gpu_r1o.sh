#!/bin/bash
# one gpurun call: LN1 folded into the FFN-in GEMM: parity + A/B bench
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_gemm.py tests/test_gpu_solve.py tests/test_gpu_configs.py -q -s > gpurun_out/t_fold.log 2>&1; echo "parity exit $?" > gpurun_out/summary.txt
i=0
for cfg in "ln1_fold=1" "ln1_fold=0" "ln1_fold=1" "ln1_fold=0"; do
  timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --opt $cfg > "gpurun_out/bench_o$i.json" 2> "gpurun_out/bench_o$i.err"; echo "bench o$i [$cfg] exit $?" >> gpurun_out/summary.txt
  i=$((i+1))
done
cat gpurun_out/summary.txt; grep -E "passed|failed|defaults|ln1_fold" gpurun_out/t_fold.log | tail -8
