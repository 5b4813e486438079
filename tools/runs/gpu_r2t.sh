#!/bin/bash
# A-operand L2 prefetch in the pair GEMM producer: same-box A/B through the option
mkdir -p gpurun_out
for o in a_l2=0 a_l2=8 a_l2=16 a_l2=0; do
  timeout 200 python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-vocoder --opt $o > gpurun_out/bench_t_$o.json 2> gpurun_out/bench_t.err; rc=$?
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_t_$o.json').read().strip().splitlines()[-1])
k=d['kernels']; print('$o', 'rc=$rc', round(d['value']), round(d['ms_per_step'],1), {m:round(k[m]['ms_per_step'],1) for m in ('gemm','gemm_qkv','gemm_attn_out','gemm_ffn_in','gemm_ffn_out','gemm_skip','attention','layernorm')}, d['clocks']['sm_mhz'])
PY
done
