#!/bin/bash
# ncu evidence for the final round-1 kernels (plain run first; short workload: 2 layers, 2 CFG passes at the configs[1] shape)
mkdir -p gpurun_out
timeout 120 python tools/profile_pass.py --layers 2 --passes 2 > gpurun_out/plain2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm2_tc -s 8 -c 9 -o gpurun_out/prof_gemm2_final -f python tools/profile_pass.py --layers 2 --passes 2 > gpurun_out/ncu_gemm2.log 2>&1
echo "ncu gemm exit $?" > gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn8 -s 2 -c 2 -o gpurun_out/prof_attn8_final -f python tools/profile_pass.py --layers 2 --passes 2 > gpurun_out/ncu_attn8.log 2>&1
echo "ncu attn exit $?" >> gpurun_out/summary.txt
timeout 120 python tools/profile_pass.py --passes 1 > gpurun_out/plain.log 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_pass.csv python tools/profile_pass.py --passes 1 > gpurun_out/ncu_list.log 2>&1
echo "launch list exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
