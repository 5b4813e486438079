#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vocoder.py tests/test_gpu_glue.py -q -m gpu -s > gpurun_out/t_voc.log 2>&1; echo "pytest vocoder+glue exit $?" > gpurun_out/summary.txt
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench default exit $?" >> gpurun_out/summary.txt
grep -E "rel-L2|bf16 vs fp32|passed|failed|Error" gpurun_out/t_voc.log | tail -12
cat gpurun_out/summary.txt; python -c "
import json; d=json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['vocoder'])"; tail -3 gpurun_out/bench_default.err
