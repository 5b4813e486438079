#!/bin/bash
# one gpurun call: first run of the vocoder path
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_vocoder.py -q -m gpu -x -s > gpurun_out/t_voc.log 2>&1; echo "pytest vocoder exit $?" > gpurun_out/summary.txt
tail -40 gpurun_out/t_voc.log
cat gpurun_out/summary.txt
