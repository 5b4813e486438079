#!/bin/bash
# one gpurun call: GEMM epilogue order restored for the non-residual epilogues (A/B against the round-1 figures), QKV norm epilogue cost
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_bf16.py tests/test_gpu_configs.py -q -m gpu -x > gpurun_out/t_h.log 2>&1; echo "pytest exit $?" > gpurun_out/summary.txt
i=0
for cfg in "qkv_norm=1" "qkv_norm=0" "attn_v=8" "qkv_norm=1" "qkv_norm=0"; do
  timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-vocoder --opt $cfg > "gpurun_out/bench_h$i.json" 2> "gpurun_out/bench_h$i.err"; echo "bench h$i [$cfg] exit $?" >> gpurun_out/summary.txt
  i=$((i+1))
done
timeout 120 python bench.py --batch 1 --seconds 5 --steps 5 --warmup 3 --no-cpu-baseline --no-vocoder --no-profile > gpurun_out/bench_b1_h.json 2>/dev/null
cat gpurun_out/summary.txt; tail -2 gpurun_out/t_h.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_h*.json'))+['gpurun_out/bench_b1_h.json']:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value']), round(d['ms_per_step'],1), d['clocks']['sm_mhz'], {k:round(v['ms_per_step'],1) for k,v in (d.get('kernels') or {}).items()})
    except Exception as e: print(f, 'ERR', e)
PY
