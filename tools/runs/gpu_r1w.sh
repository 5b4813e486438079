#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_bf16.py -q -m gpu -s > gpurun_out/t_attn.log 2>&1; echo "pytest exit $?" > gpurun_out/summary.txt
timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-vocoder > gpurun_out/bench_w0.json 2> gpurun_out/bench_w0.err; echo "bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; grep -E "spike|defaults|passed|failed|Error" gpurun_out/t_attn.log | tail -9
python -c "
import json; d=json.loads(open('gpurun_out/bench_w0.json').read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],1), {k:round(v['ms_per_step'],1) for k,v in d['kernels'].items() if not k.startswith('gemm_')})"
