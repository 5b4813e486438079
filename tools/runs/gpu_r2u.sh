#!/bin/bash
mkdir -p gpurun_out
for o in gemm_mc=0 gemm_mc=1; do
  timeout 200 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-vocoder --opt $o > gpurun_out/bench_u_$o.json 2> gpurun_out/bench_u.err; rc=$?
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_u_$o.json').read().strip().splitlines()[-1])
k=d['kernels']; print('$o', 'rc=$rc', round(d['value']), round(d['ms_per_step'],1), {m:(round(k[m]['ms_per_step'],1),k[m]['launches']) for m in ('gemm','gemm_qkv','gemm_attn_out','gemm_ffn_in','gemm_ffn_out','gemm_skip','attention','layernorm')}, d['clocks']['sm_mhz'])
PY
done
