#!/bin/bash
# one gpurun call: attention v10 with the cost-sorted item order: kernel tests, lab, bench A/B (sort on/off, poly 3/4)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_attention.py -q -m gpu -x -k "10- or 100- or v9 or long_form or persistent" > gpurun_out/t_attn10b.log 2>&1; echo "pytest attention v10 exit $?" > gpurun_out/summary.txt
timeout 300 ./tools/attn_lab 0.45 1134 32 > gpurun_out/attn_lab4.txt 2>&1; echo "attn_lab exit $?" >> gpurun_out/summary.txt
timeout 200 ./tools/attn_lab 0.45 2856 16 > gpurun_out/attn_lab4_long.txt 2>&1; echo "attn_lab long exit $?" >> gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_configs.py tests/test_gpu_solve.py -q -m gpu -x > gpurun_out/t_main.log 2>&1; echo "pytest configs/solve exit $?" >> gpurun_out/summary.txt
i=0
for cfg in "attn_sort=1" "attn_sort=0" "attn_poly=3" "attn_sort=1" "attn_sort=0" "attn_poly=3"; do
  timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-vocoder --opt $cfg > "gpurun_out/bench_e$i.json" 2> "gpurun_out/bench_e$i.err"; echo "bench e$i [$cfg] exit $?" >> gpurun_out/summary.txt
  i=$((i+1))
done
cat gpurun_out/summary.txt; tail -2 gpurun_out/t_attn10b.log; tail -2 gpurun_out/t_main.log; grep -E "window=|v10|plan" gpurun_out/attn_lab4.txt | head -30; grep -E "window=|v10 win=. poly=3 lab= 0|plan" gpurun_out/attn_lab4_long.txt
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_e*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value']), round(d['ms_per_step'],1), d['clocks']['sm_mhz'], {k:round(v['ms_per_step'],1) for k,v in (d.get('kernels') or {}).items() if not k.startswith('gemm_')})
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-600:])
PY
