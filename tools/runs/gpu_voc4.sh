#!/bin/bash
# one gpurun call: full GPU suite + smoke + vocoder ncu evidence (plain run first)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/t_all.log 2>&1; echo "pytest -m gpu exit $?" > gpurun_out/summary.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt
timeout 120 python tools/profile_vocoder.py > gpurun_out/voc_plain.log 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_vocoder.csv python tools/profile_vocoder.py > gpurun_out/ncu_voc_list.log 2>&1
echo "vocoder launch list exit $?" >> gpurun_out/summary.txt
# stage-1 conv (k=11, C=768), stage-6 conv and one activation of a middle stage, full sections
timeout 600 ncu --set full --clock-control none --import-source on -k regex:voc_conv_tc -s 20 -c 3 -o gpurun_out/prof_voc_conv_wide -f python tools/profile_vocoder.py > gpurun_out/ncu_voc1.log 2>&1
echo "ncu conv wide exit $?" >> gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:voc_conv_tc -s 100 -c 3 -o gpurun_out/prof_voc_conv_narrow -f python tools/profile_vocoder.py > gpurun_out/ncu_voc2.log 2>&1
echo "ncu conv narrow exit $?" >> gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_voc_act -s 40 -c 2 -o gpurun_out/prof_voc_act -f python tools/profile_vocoder.py > gpurun_out/ncu_voc3.log 2>&1
echo "ncu act exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -3 gpurun_out/t_all.log; tail -4 gpurun_out/smoke.log
