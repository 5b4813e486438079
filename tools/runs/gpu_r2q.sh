#!/bin/bash
# single-utterance (B = 1, 5 s + prompt) option sweep on one box
mkdir -p gpurun_out
for o in graph=1 ln_fuse=2 attn_v=8 attn_sort=0 attn_win=0 gemm_bn=0 pdl=0 zigzag=0 ln1_fold=0; do
  timeout 200 python bench.py --batch 1 --seconds 5 --steps 5 --warmup 3 --no-cpu-baseline --no-vocoder --no-profile --opt $o > gpurun_out/bench_b1_$o.json 2> gpurun_out/bench_q.err; rc=$?
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_b1_$o.json').read().strip().splitlines()[-1])
    print('$o', 'rc=$rc', round(d['value']), round(d['ms_per_step'],2), d['gpu_launches'], d['clocks']['sm_mhz'])
except Exception as e: print('$o', 'rc=$rc', 'failed', e)
PY
done
