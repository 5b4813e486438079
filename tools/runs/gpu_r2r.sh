#!/bin/bash
# packed fp32x2 GELU in the FFN-in epilogue: tests (bit-level properties incl.) + same-box A/B against the previous build
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_bf16.py tests/test_gpu_configs.py -x -q -m gpu > gpurun_out/pytest_r2r.log 2>&1; echo "pytest exit $?"
tail -3 gpurun_out/pytest_r2r.log
for rep in 1 2; do
  for lib in prev cur; do
    if [ $lib = prev ]; then export TVB_LIBRARY=$PWD/usdm_b200/csrc/libtvb_prev.so; else unset TVB_LIBRARY; fi
    timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-vocoder > gpurun_out/bench_gelu2_${lib}_$rep.json 2> gpurun_out/bench_r.err; echo "bench $lib $rep exit $?"
  done
done
unset TVB_LIBRARY
python - <<'PY'
import json
for n in ('prev_1','cur_1','prev_2','cur_2'):
    try:
        d=json.loads(open(f'gpurun_out/bench_gelu2_{n}.json').read().strip().splitlines()[-1])
        k=d['kernels']; print(n, round(d['value']), round(d['ms_per_step'],1), {m:round(k[m]['ms_per_step'],1) for m in ('gemm','gemm_qkv','gemm_attn_out','gemm_ffn_in','gemm_ffn_out','gemm_skip','attention','layernorm')}, d['clocks']['sm_mhz'])
    except Exception as e: print(n,'failed',e)
PY
