#!/bin/bash
# ncu --set full of the GEMM family with the final kernels (plain run first)
mkdir -p gpurun_out
timeout 120 python tools/profile_pass.py --layers 2 --passes 2 > gpurun_out/plain2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm2_tc -s 10 -c 10 -o gpurun_out/prof_gemm2_final -f python tools/profile_pass.py --layers 2 --passes 2 > gpurun_out/ncu_gemm2.log 2>&1
echo "ncu gemm exit $?" > gpurun_out/summary.txt
cat gpurun_out/summary.txt
