#!/bin/bash
# one gpurun call: attention v10 (whole additive term from the tensor core): kernel tests, lab timings, bench A/B against v9
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_attention.py -q -m gpu -x -s -k "v9 or 10- or 100- or long_form or persistent" > gpurun_out/t_attn10.log 2>&1; echo "pytest attention v10 exit $?" > gpurun_out/summary.txt
timeout 300 ./tools/attn_lab 0.45 1134 32 > gpurun_out/attn_lab3.txt 2>&1; echo "attn_lab exit $?" >> gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_configs.py tests/test_gpu_solve.py -q -m gpu -x > gpurun_out/t_main.log 2>&1; echo "pytest bf16/configs/solve exit $?" >> gpurun_out/summary.txt
i=0
for cfg in "attn_v=10" "attn_v=9" "attn_v=10 --opt resid_l2=2" "attn_v=10" "attn_v=9"; do
  timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-vocoder --opt $cfg > "gpurun_out/bench_d$i.json" 2> "gpurun_out/bench_d$i.err"; echo "bench d$i [$cfg] exit $?" >> gpurun_out/summary.txt
  i=$((i+1))
done
cat gpurun_out/summary.txt; tail -4 gpurun_out/t_attn10.log; tail -3 gpurun_out/t_main.log; head -40 gpurun_out/attn_lab3.txt
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_d*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value']), round(d['ms_per_step'],1), d['clocks']['sm_mhz'], {k:round(v['ms_per_step'],1) for k,v in (d.get('kernels') or {}).items()})
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-600:])
PY
