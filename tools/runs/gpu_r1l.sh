#!/bin/bash
# one gpurun call: ncu evidence for the shipping defaults (plain run first, then the launch list and two --set full captures)
mkdir -p gpurun_out
timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 21700 -c 7300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
echo "launch list exit $?" > gpurun_out/summary.txt
timeout 120 python tools/profile_pass.py --layers 2 --passes 2 > gpurun_out/plain2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm2_tc -s 8 -c 8 -o gpurun_out/prof_gemm2 -f python tools/profile_pass.py --layers 2 --passes 2 > gpurun_out/ncu_gemm2.log 2>&1
echo "ncu gemm exit $?" >> gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn8 -s 2 -c 2 -o gpurun_out/prof_attn8 -f python tools/profile_pass.py --layers 2 --passes 2 > gpurun_out/ncu_attn8.log 2>&1
echo "ncu attn exit $?" >> gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none -k regex:k_layernorm -s 4 -c 2 -o gpurun_out/prof_ln -f python tools/profile_pass.py --layers 2 --passes 2 > gpurun_out/ncu_ln.log 2>&1
echo "ncu ln exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
