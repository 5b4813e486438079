#!/bin/bash
# final validation of the committed tree: GPU tests, smoke, default bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest -m gpu exit $?"
tail -2 gpurun_out/pytest_gpu_final.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_final.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke_final.log
timeout 600 python bench.py > gpurun_out/bench_final_check.json 2> gpurun_out/bench_final_check.err; echo "bench exit $?"; cut -c1-260 gpurun_out/bench_final_check.json
