#!/bin/bash
# one gpurun call: ncu evidence for round 2 (plain runs first): attention v10 --set full with source, launch list of one pass, GEMM --set full; B = 1 latency A/B
mkdir -p gpurun_out
for cfg in "pdl=1" "attn_v=8"; do
  timeout 120 python bench.py --batch 1 --seconds 5 --steps 5 --warmup 3 --no-cpu-baseline --no-vocoder --no-profile --opt $cfg > "gpurun_out/bench_b1_$cfg.json" 2>/dev/null; echo "bench B=1 [$cfg] exit $?" >> gpurun_out/summary.txt
done
timeout 120 python tools/profile_pass.py --layers 2 --passes 2 > gpurun_out/plain2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn10_kernel -s 2 -c 2 -o gpurun_out/prof_attn10 -f python tools/profile_pass.py --layers 2 --passes 2 > gpurun_out/ncu_attn10.log 2>&1
echo "ncu attn10 exit $?" >> gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm2_tc -s 8 -c 9 -o gpurun_out/prof_gemm2_r2 -f python tools/profile_pass.py --layers 2 --passes 2 > gpurun_out/ncu_gemm2.log 2>&1
echo "ncu gemm exit $?" >> gpurun_out/summary.txt
timeout 120 python tools/profile_pass.py --passes 1 > gpurun_out/plain.log 2>&1 && \
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_pass_r2.csv python tools/profile_pass.py --passes 1 > gpurun_out/ncu_list.log 2>&1
echo "launch list exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_b1_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value']), round(d['ms_per_step'],2), d['gpu_launches'])
    except Exception as e: print(f, 'ERR', e)
PY
