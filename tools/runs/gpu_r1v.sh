#!/bin/bash
# one gpurun call: attention sum-guard: attention / parity suites + bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_bf16.py tests/test_gpu_configs.py -q -m gpu -x -s > gpurun_out/t_attn.log 2>&1; echo "pytest exit $?" > gpurun_out/summary.txt
for i in 0 1; do
  timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-vocoder > "gpurun_out/bench_v$i.json" 2> "gpurun_out/bench_v$i.err"; echo "bench v$i exit $?" >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt; grep -E "defaults|passed|failed|Error" gpurun_out/t_attn.log | tail -6
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_v*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value']), round(d['ms_per_step'],1), d['clocks']['sm_mhz'], {k:round(v['ms_per_step'],1) for k,v in (d.get('kernels') or {}).items() if not k.startswith('gemm_')})
    except Exception as e: print(f, 'ERR', e)
PY
