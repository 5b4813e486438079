#!/bin/bash
# one gpurun call: full GPU suite with the shipping defaults (attn v8 poly 3, try_wait suspend hint), smoke, default bench with cpu baseline
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/t_all.log 2>&1; echo "pytest -m gpu exit $?" > gpurun_out/summary.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt
timeout 300 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench default exit $?" >> gpurun_out/summary.txt
timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --opt attn_poly=2 > gpurun_out/bench_k1.json 2> gpurun_out/bench_k1.err; echo "bench poly2 exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
