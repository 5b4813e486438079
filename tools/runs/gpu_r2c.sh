#!/bin/bash
# one gpurun call: residual L2 prefetch of the pair GEMM epilogue (A/B) + attention exp2 polynomial share (A/B)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_bf16.py tests/test_gpu_configs.py -q -m gpu -x > gpurun_out/t_c.log 2>&1; echo "pytest exit $?" > gpurun_out/summary.txt
i=0
for cfg in "resid_l2=1" "resid_l2=0" "resid_l2=1 --opt attn_poly=4" "resid_l2=1" "resid_l2=0" "resid_l2=1 --opt attn_poly=4"; do
  timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-vocoder --opt $cfg > "gpurun_out/bench_c$i.json" 2> "gpurun_out/bench_c$i.err"; echo "bench c$i [$cfg] exit $?" >> gpurun_out/summary.txt
  i=$((i+1))
done
cat gpurun_out/summary.txt; tail -3 gpurun_out/t_c.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_c*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value']), round(d['ms_per_step'],1), d['clocks']['sm_mhz'], {k:round(v['ms_per_step'],1) for k,v in (d.get('kernels') or {}).items()})
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-600:])
PY
