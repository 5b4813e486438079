#!/bin/bash
# GEMM epilogue: per-column constants requested ahead of the waits; FOLD partials one tile ahead.  Tests + bench A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_bf16.py tests/test_gpu_configs.py -x -q -m gpu > gpurun_out/pytest_r2l.log 2>&1; echo "pytest exit $?" > gpurun_out/summary.txt
tail -3 gpurun_out/pytest_r2l.log
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-vocoder > gpurun_out/bench_r2l.json 2> gpurun_out/bench_r2l.err; echo "bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2l.json').read().strip().splitlines()[-1])
k=d['kernels']; print(round(d['value']), round(d['ms_per_step'],1), {n:round(k[n]['ms_per_step'],1) for n in k}, d['clocks'])
PY
