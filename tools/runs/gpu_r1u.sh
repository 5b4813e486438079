#!/bin/bash
# one gpurun call: tensor-core frame projection: parity suites + A/B bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_solve.py tests/test_gpu_configs.py tests/test_gpu_fp32.py -q -m gpu -x -s > gpurun_out/t_frame.log 2>&1; echo "pytest exit $?" > gpurun_out/summary.txt
i=0
for cfg in "conv_halo=1" "conv_halo=0" "conv_halo=1" "conv_halo=0"; do
  timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-vocoder --opt $cfg > "gpurun_out/bench_u$i.json" 2> "gpurun_out/bench_u$i.err"; echo "bench u$i [$cfg] exit $?" >> gpurun_out/summary.txt
  i=$((i+1))
done
timeout 120 python bench.py --batch 1 --seconds 5 --steps 5 --warmup 3 --no-cpu-baseline --no-vocoder --no-profile > gpurun_out/bench_ub1.json 2>/dev/null; echo "bench B=1 exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; grep -E "rel|passed|failed|Error" gpurun_out/t_frame.log | tail -15
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_u*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value']), round(d['ms_per_step'],1), d['clocks']['sm_mhz'], {k:round(v['ms_per_step'],1) for k,v in (d.get('kernels') or {}).items() if not k.startswith('gemm_')})
    except Exception as e: print(f, 'ERR', e)
PY
