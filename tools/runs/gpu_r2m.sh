#!/bin/bash
# same-box A/B: previous commit's library (libtvb_prev.so) against the working tree's, alternating; then the no-traffic floor of attention-out
mkdir -p gpurun_out
for rep in 1 2; do
  for lib in prev cur; do
    if [ $lib = prev ]; then export TVB_LIBRARY=$PWD/usdm_b200/csrc/libtvb_prev.so; else unset TVB_LIBRARY; fi
    timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-vocoder > gpurun_out/bench_ab_${lib}_$rep.json 2> gpurun_out/bench_ab_${lib}_$rep.err; echo "bench $lib $rep exit $?" >> gpurun_out/summary_m.txt
  done
done
unset TVB_LIBRARY
timeout 200 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-vocoder --opt gemm_lab=7 > gpurun_out/bench_ab_cur_lab7.json 2>/dev/null
cat gpurun_out/summary_m.txt
python - <<'PY'
import json
for n in ('prev_1','cur_1','prev_2','cur_2','cur_lab7'):
    try:
        d=json.loads(open(f'gpurun_out/bench_ab_{n}.json').read().strip().splitlines()[-1])
        k=d['kernels']; print(n, round(d['value']), round(d['ms_per_step'],1), {m:round(k[m]['ms_per_step'],1) for m in ('gemm','gemm_qkv','gemm_attn_out','gemm_ffn_in','gemm_ffn_out','gemm_skip','attention','layernorm')}, d['clocks']['sm_mhz'])
    except Exception as e: print(n,'failed',e)
PY
