#!/bin/bash
mkdir -p gpurun_out
timeout 90 ./tools/attn_lab > gpurun_out/attn_lab_v12.txt 2>&1; echo "lab exit $?"
grep -v "role" gpurun_out/attn_lab_v12.txt | grep "v12\|v10 win=. poly=3 lab= 0\|v10 win=. poly=4 lab= 0\|v10 win=0 poly=3 lab=11\|error\|rror"
