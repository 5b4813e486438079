#!/bin/bash
# one gpurun call: single-request (B = 1) breakdown + units -> waveform latency
mkdir -p gpurun_out
timeout 200 python bench.py --batch 1 --seconds 5 --steps 5 --warmup 3 --no-cpu-baseline --no-vocoder > gpurun_out/bench_b1_prof.json 2> gpurun_out/bench_b1_prof.err; echo "bench B=1 exit $?" > gpurun_out/summary.txt
timeout 300 python tools/latency.py > gpurun_out/latency.jsonl 2> gpurun_out/latency.err; echo "latency exit $?" >> gpurun_out/summary.txt
timeout 120 python tools/profile_pass.py --batch 1 --frames 702 --passes 3 > gpurun_out/plain_b1.log 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gemm|attn|conv|k_layernorm|k_pack_y|k_ode|k_frame|k_scatter|k_sgemm|k_fill' -s 330 -c 700 --csv --log-file gpurun_out/launches_b1.csv python tools/profile_pass.py --batch 1 --frames 702 --passes 3 > gpurun_out/ncu_b1.log 2>&1
echo "launch list B=1 exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; cat gpurun_out/latency.jsonl
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_b1_prof.json').read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],2), d['gpu_launches'], {k:(round(v['ms_per_step'],2),v['launches']) for k,v in (d.get('kernels') or {}).items()})
PY
