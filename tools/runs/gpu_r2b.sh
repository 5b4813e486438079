#!/bin/bash
# one gpurun call: attention lab (timing ablations of attn9) + ncu --set full capture of attn9 with source
mkdir -p gpurun_out
timeout 300 ./tools/attn_lab 0.45 1134 32 > gpurun_out/attn_lab.txt 2>&1; echo "attn_lab exit $?" > gpurun_out/summary.txt
timeout 120 python tools/profile_pass.py --layers 2 --passes 2 > gpurun_out/plain2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn9_kernel -s 2 -c 2 -o gpurun_out/prof_attn9 -f python tools/profile_pass.py --layers 2 --passes 2 > gpurun_out/ncu_attn9.log 2>&1
echo "ncu attn9 exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; cat gpurun_out/attn_lab.txt; tail -3 gpurun_out/ncu_attn9.log
