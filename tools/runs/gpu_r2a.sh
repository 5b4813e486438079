#!/bin/bash
# one gpurun call (round 2, first): pipe microbenchmarks, the residual-staging fix of the pair GEMM, attention v9 (fixed reference + ALiBi window)
mkdir -p gpurun_out
timeout 120 ./tools/pipe_bench > gpurun_out/pipe_bench.txt 2>&1; echo "pipe_bench exit $?" > gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_attention.py -q -m gpu -x -s -k "v9 or 9- or 90- or long_form or persistent" > gpurun_out/t_attn9.log 2>&1; echo "pytest attention v9 exit $?" >> gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_bf16.py tests/test_gpu_configs.py tests/test_gpu_solve.py -q -m gpu -x -s > gpurun_out/t_main.log 2>&1; echo "pytest gemm/bf16/configs/solve exit $?" >> gpurun_out/summary.txt
i=0
for cfg in "attn_v=9" "attn_v=8" "attn_v=9 --opt attn_win=0" "attn_v=9" "attn_v=8"; do
  timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-vocoder --opt $cfg > "gpurun_out/bench_a$i.json" 2> "gpurun_out/bench_a$i.err"; echo "bench a$i [$cfg] exit $?" >> gpurun_out/summary.txt
  i=$((i+1))
done
cat gpurun_out/summary.txt; tail -5 gpurun_out/t_attn9.log; tail -5 gpurun_out/t_main.log; cat gpurun_out/pipe_bench.txt
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_a*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value']), round(d['ms_per_step'],1), d['clocks']['sm_mhz'], {k:round(v['ms_per_step'],1) for k,v in (d.get('kernels') or {}).items()})
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-600:])
PY
