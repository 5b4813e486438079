#!/bin/bash
# one gpurun call: 5-deep vs 4-deep operand ring in the non-ST GEMM epilogues (two prebuilt libraries swapped in place)
mkdir -p gpurun_out
cd usdm_b200/csrc; cp libtvb.so libtvb_s5.so; cd ../..
timeout 300 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_bf16.py -q -m gpu -x > gpurun_out/t_gemm.log 2>&1; echo "pytest gemm+bf16 (5 stages) exit $?" > gpurun_out/summary.txt
i=0
for v in s5 s4 s5 s4; do
  cp usdm_b200/csrc/libtvb_$v.so usdm_b200/csrc/libtvb.so
  timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > "gpurun_out/bench_t${i}_$v.json" 2> "gpurun_out/bench_t$i.err"; echo "bench t$i [$v] exit $?" >> gpurun_out/summary.txt
  i=$((i+1))
done
cat gpurun_out/summary.txt; tail -3 gpurun_out/t_gemm.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_t*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value']), round(d['ms_per_step'],1), d['clocks']['sm_mhz'], {k:round(v['ms_per_step'],1) for k,v in (d.get('kernels') or {}).items()})
    except Exception as e: print(f, 'ERR', e)
PY
