#!/bin/bash
# residual L2 prefetch re-measured on the current epilogue (same box)
mkdir -p gpurun_out
for o in resid_l2=0 resid_l2=1 resid_l2=2 resid_l2=0; do
  timeout 200 python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-vocoder --opt $o > gpurun_out/bench_p_$o.json 2> gpurun_out/bench_p.err; echo "bench $o exit $?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_p_$o.json').read().strip().splitlines()[-1])
k=d['kernels']; print('$o', round(d['value']), round(d['ms_per_step'],1), {m:round(k[m]['ms_per_step'],1) for m in ('gemm','gemm_qkv','gemm_attn_out','gemm_ffn_in','gemm_ffn_out','gemm_skip','attention','layernorm')}, d['clocks']['sm_mhz'])
PY
done
