#!/bin/bash
# one gpurun call: programmatic dependent launch: full suite + A/B bench + B=1 latency
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/t_all.log 2>&1; echo "pytest -m gpu exit $?" > gpurun_out/summary.txt
i=0
for cfg in "pdl=1" "pdl=0" "pdl=1" "pdl=0"; do
  timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-profile --opt $cfg > "gpurun_out/bench_s$i.json" 2> "gpurun_out/bench_s$i.err"; echo "bench s$i [$cfg] exit $?" >> gpurun_out/summary.txt
  i=$((i+1))
done
for cfg in "pdl=1" "pdl=0"; do
  timeout 120 python bench.py --batch 1 --seconds 5 --steps 5 --warmup 3 --no-cpu-baseline --no-profile --opt $cfg > "gpurun_out/bench_b1_$cfg.json" 2>/dev/null; echo "bench B=1 [$cfg] exit $?" >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt; tail -3 gpurun_out/t_all.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_s*.json'))+sorted(glob.glob('gpurun_out/bench_b1_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value']), round(d['ms_per_step'],1), d['clocks']['sm_mhz'])
    except Exception as e: print(f, 'ERR', e)
PY
