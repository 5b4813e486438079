#!/bin/bash
# Round-2 final evidence, one GPU: full GPU test suite, smoke, default bench (+ reference arm), sweep, latency, ncu launch list of bench.py
# itself (a window inside the timed step) and ncu --set full captures of the GEMM family and the attention kernel (plain runs first).
mkdir -p gpurun_out
S=gpurun_out/summary_s.txt; : > $S
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_all.log 2>&1; echo "pytest -m gpu exit $?" >> $S
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> $S
timeout 600 python bench.py > gpurun_out/bench_default_final.json 2> gpurun_out/bench_default_final.err; echo "bench default exit $?" >> $S
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_reference_cpu.json 2> gpurun_out/bench_reference_cpu.err; echo "bench reference exit $?" >> $S
timeout 600 python tools/sweep.py > gpurun_out/sweep_1gpu.jsonl 2> gpurun_out/sweep_1gpu.err; echo "sweep exit $?" >> $S
timeout 300 python tools/latency.py > gpurun_out/latency.jsonl 2> gpurun_out/latency.err; echo "latency exit $?" >> $S
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 22000 -c 700 --csv --log-file gpurun_out/ncu_launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-vocoder --no-profile > gpurun_out/ncu_bench.log 2>&1; echo "ncu launch list (bench.py) exit $?" >> $S
timeout 120 python tools/profile_pass.py --layers 2 --passes 2 > gpurun_out/plain2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm2_tc -s 8 -c 9 -o gpurun_out/prof_gemm2_r2f -f python tools/profile_pass.py --layers 2 --passes 2 > gpurun_out/ncu_gemm2.log 2>&1
echo "ncu gemm exit $?" >> $S
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn10_kernel -s 2 -c 2 -o gpurun_out/prof_attn10_r2f -f python tools/profile_pass.py --layers 2 --passes 2 > gpurun_out/ncu_attn10.log 2>&1
echo "ncu attn10 exit $?" >> $S
cat $S; tail -2 gpurun_out/pytest_gpu_all.log; tail -5 gpurun_out/smoke.log; cut -c1-400 gpurun_out/bench_default_final.json; cut -c1-250 gpurun_out/sweep_1gpu.jsonl; cat gpurun_out/latency.jsonl | cut -c1-300
