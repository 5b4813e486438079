#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_bf16.py -x -q -s -k "ln_fuse" > gpurun_out/t_lnf.log 2>&1; rc=$?; echo "ln_fuse=2 parity exit $rc" > gpurun_out/summary.txt
if [ $rc -eq 0 ]; then
  timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --opt ln_fuse=2 > gpurun_out/bench_m0.json 2> gpurun_out/bench_m0.err; echo "bench ln_fuse=2 exit $?" >> gpurun_out/summary.txt
fi
timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_m1.json 2> gpurun_out/bench_m1.err; echo "bench default exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -5 gpurun_out/t_lnf.log
