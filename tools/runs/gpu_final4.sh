#!/bin/bash
# one gpurun call: final round-1 evidence with the shipping defaults (plain runs first, ncu after)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/t_all.log 2>&1; echo "pytest -m gpu exit $?" > gpurun_out/summary.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench default exit $?" >> gpurun_out/summary.txt
timeout 600 python tools/sweep.py > gpurun_out/sweep_1gpu.jsonl 2> gpurun_out/sweep.err; echo "sweep exit $?" >> gpurun_out/summary.txt
timeout 120 python tools/profile_pass.py --passes 1 > gpurun_out/plain.log 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gemm|attn|conv|k_layernorm|k_pack_y|k_ode|k_frame|k_scatter|k_sgemm' -c 1200 --csv --log-file gpurun_out/launches_pass.csv python tools/profile_pass.py --passes 1 > gpurun_out/ncu_list.log 2>&1
echo "launch list exit $?" >> gpurun_out/summary.txt
timeout 120 python tools/profile_pass.py --layers 2 --passes 2 > gpurun_out/plain2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn8 -s 2 -c 2 -o gpurun_out/prof_attn8_final -f python tools/profile_pass.py --layers 2 --passes 2 > gpurun_out/ncu_attn8.log 2>&1
echo "ncu attn exit $?" >> gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_halo_kernel -s 2 -c 2 -o gpurun_out/prof_conv_halo -f python tools/profile_pass.py --layers 2 --passes 2 > gpurun_out/ncu_convh.log 2>&1
echo "ncu conv halo exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -3 gpurun_out/t_all.log; tail -4 gpurun_out/smoke.log; cat gpurun_out/sweep_1gpu.jsonl | cut -c1-250
