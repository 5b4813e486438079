#!/bin/bash
# attention v11 (64-key pieces) in the lab: timing + ablations + bitwise comparison with v10 + trace
mkdir -p gpurun_out
timeout 300 ./tools/attn_lab > gpurun_out/attn_lab_v11.txt 2>&1; echo "lab exit $?" > gpurun_out/summary.txt
cat gpurun_out/summary.txt; grep -v "role" gpurun_out/attn_lab_v11.txt | head -70
