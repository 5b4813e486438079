#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_vocoder.py -q -m gpu -s > gpurun_out/t_voc.log 2>&1; echo "pytest vocoder exit $?" > gpurun_out/summary.txt
timeout 300 python tools/bench_vocoder.py --cpu-baseline > gpurun_out/bench_voc.json 2> gpurun_out/bench_voc.err; echo "bench vocoder exit $?" >> gpurun_out/summary.txt
timeout 300 python tools/bench_vocoder.py --batch 1 --seconds 5 > gpurun_out/bench_voc_b1.json 2>> gpurun_out/bench_voc.err; echo "bench vocoder b1 exit $?" >> gpurun_out/summary.txt
grep -E "rel-L2|passed|failed|Error|error" gpurun_out/t_voc.log | tail -30
cat gpurun_out/summary.txt; cat gpurun_out/bench_voc.json gpurun_out/bench_voc_b1.json; tail -5 gpurun_out/bench_voc.err
