#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench default exit $?" > gpurun_out/summary.txt
cat gpurun_out/summary.txt; python -c "
import json; d=json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['vocoder'])"; tail -3 gpurun_out/bench_default.err
