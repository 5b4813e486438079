#!/bin/bash
mkdir -p gpurun_out
for v in "" _spin1 _spin2; do
timeout 90 ./tools/attn_lab$v > gpurun_out/attn_lab_wait$v.txt 2>&1; echo "lab$v exit $?"
grep -v "role" gpurun_out/attn_lab_wait$v.txt | grep "v10 win=. poly=3 lab= 0\|v10 win=. poly=4 lab= 0\|v10 win=0 poly=3 lab=11\|v11 win=. poly=3 lab= 0\|v11 win=0 poly=3 lab=11\|rror"
done
