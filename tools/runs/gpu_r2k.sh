#!/bin/bash
# (1) attention v11 lab with the unsplit-QK variant; (2) timing ablations of the attention-out residual epilogue (gemm_lab bits)
mkdir -p gpurun_out
timeout 300 ./tools/attn_lab > gpurun_out/attn_lab_v11b.txt 2>&1; echo "lab exit $?" > gpurun_out/summary.txt
grep -v "role" gpurun_out/attn_lab_v11b.txt | grep "v11\|v10 win=. poly=3 lab= 0\|v10 win=. poly=4 lab= 0" 
for lab in 0 1 2 4 3 7; do
  timeout 200 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-vocoder --opt gemm_lab=$lab > gpurun_out/bench_gemm_lab$lab.json 2> gpurun_out/bench_gemm_lab$lab.err; echo "bench gemm_lab=$lab exit $?" >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt
python - <<'PY'
import json
for lab in (0,1,2,4,3,7):
    try:
        d=json.loads(open(f'gpurun_out/bench_gemm_lab{lab}.json').read().strip().splitlines()[-1])
        k=d['kernels']; print(lab, round(d['ms_per_step'],1), {n:round(k[n]['ms_per_step'],1) for n in ('gemm_attn_out','gemm_ffn_in','gemm_ffn_out','gemm_qkv','gemm_skip','attention','layernorm')}, d['clocks']['sm_mhz'])
    except Exception as e: print(lab, 'failed', e)
PY
