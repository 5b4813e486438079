#!/bin/bash
# one gpurun call: LayerNorm fused into the residual GEMM epilogue (8-CTA clusters): parity + bench; launch list of a short pass
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_bf16.py -x -q -s -k "ln_fuse=2" > gpurun_out/t_lnf.log 2>&1; rc=$?; echo "ln_fuse=2 parity exit $rc" > gpurun_out/summary.txt
if [ $rc -eq 0 ]; then
  timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --opt ln_fuse=2 > gpurun_out/bench_m0.json 2> gpurun_out/bench_m0.err; echo "bench ln_fuse=2 exit $?" >> gpurun_out/summary.txt
fi
timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_m1.json 2> gpurun_out/bench_m1.err; echo "bench default exit $?" >> gpurun_out/summary.txt
timeout 120 python tools/profile_pass.py --passes 1 > gpurun_out/plain.log 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_pass.csv python tools/profile_pass.py --passes 1 > gpurun_out/ncu_list.log 2>&1
echo "launch list exit $?" >> gpurun_out/summary.txt
timeout 300 python -m pytest tests/test_gpu_solve.py tests/test_gpu_configs.py -q > gpurun_out/t_solve.log 2>&1; echo "solve+configs exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -5 gpurun_out/t_lnf.log
