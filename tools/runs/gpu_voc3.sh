#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_vocoder.py -q -s -k "ablation or golden" > gpurun_out/halo_check.log 2>&1; cat gpurun_out/halo_check.log
timeout 600 python -m pytest tests/test_gpu_vocoder.py -q -m gpu > gpurun_out/t_voc.log 2>&1; echo "pytest vocoder exit $?" > gpurun_out/summary.txt
for hl in 1 0; do timeout 300 python tools/bench_vocoder.py --opt act_vec2=$hl > gpurun_out/bench_voc_h$hl.json 2>> gpurun_out/bench_voc.err; echo "bench vocoder act_vec2=$hl exit $?" >> gpurun_out/summary.txt; done
tail -3 gpurun_out/t_voc.log
cat gpurun_out/summary.txt; cat gpurun_out/bench_voc_h1.json gpurun_out/bench_voc_h0.json; tail -5 gpurun_out/bench_voc.err
