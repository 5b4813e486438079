#!/bin/bash
# one gpurun call: checkpoint of round 2 with the shipping defaults (attention v10): full GPU suite, smoke, bench, sweep over BASELINE configs
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/t_all.log 2>&1; echo "pytest -m gpu exit $?" > gpurun_out/summary.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench default exit $?" >> gpurun_out/summary.txt
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench reference exit $?" >> gpurun_out/summary.txt
timeout 600 python tools/sweep.py > gpurun_out/sweep_1gpu.jsonl 2> gpurun_out/sweep.err; echo "sweep exit $?" >> gpurun_out/summary.txt
for cfg in "attn_sort=0" "attn_sort=2"; do
  timeout 600 python tools/sweep.py --opt $cfg > "gpurun_out/sweep_1gpu_$cfg.jsonl" 2>> gpurun_out/sweep.err; echo "sweep [$cfg] exit $?" >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt; tail -3 gpurun_out/t_all.log; tail -4 gpurun_out/smoke.log; cut -c1-230 gpurun_out/sweep_1gpu*.jsonl; cut -c1-400 gpurun_out/bench_default.json; cut -c1-300 gpurun_out/bench_reference.json
