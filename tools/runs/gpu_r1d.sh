#!/bin/bash
# one gpurun call: cluster-of-4 multicast GEMM: unit test (short timeout: a protocol error would hang), parity, bench
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_gemm.py -x -q -k "cluster4" > gpurun_out/t_gemm4.log 2>&1; rc=$?; echo "gemm4 exit $rc" > gpurun_out/summary.txt
if [ $rc -ne 0 ]; then
  echo "multicast GEMM failed; benching with gemm_mc=0" >> gpurun_out/summary.txt
  timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --opt gemm_mc=0 --opt attn_poly=3 > "gpurun_out/bench_mc0.json" 2> "gpurun_out/bench_mc0.err"; echo "bench mc0 exit $?" >> gpurun_out/summary.txt
  cat gpurun_out/summary.txt; exit 0
fi
timeout 300 python -m pytest tests/test_gpu_gemm.py -x -q > gpurun_out/t_gemm.log 2>&1; echo "gemm exit $?" >> gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_bf16.py -q -s > gpurun_out/t_bf16.log 2>&1; echo "bf16 exit $?" >> gpurun_out/summary.txt
for cfg in "attn_poly=3" "gemm_mc=0"; do
  timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --opt $cfg > "gpurun_out/bench_${cfg}.json" 2> "gpurun_out/bench_${cfg}.err"; echo "bench $cfg exit $?" >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt
