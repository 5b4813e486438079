#!/bin/bash
# one gpurun call: CUDA-graph replay for small batches: solve tests + latency sweep
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_solve.py tests/test_gpu_configs.py tests/test_gpu_glue.py -q -s > gpurun_out/t_graph.log 2>&1; echo "solve tests exit $?" > gpurun_out/summary.txt
timeout 600 python tools/sweep.py > gpurun_out/sweep_1gpu.jsonl 2> gpurun_out/sweep_1gpu.err; echo "sweep exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; grep -E "passed|failed|Error" gpurun_out/t_graph.log | tail -5; cat gpurun_out/sweep_1gpu.jsonl | cut -c1-260
