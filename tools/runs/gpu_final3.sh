#!/bin/bash
# one gpurun call: final round-1 evidence with the shipping defaults (plain runs first, ncu after)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/t_all.log 2>&1; echo "pytest -m gpu exit $?" > gpurun_out/summary.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench default exit $?" >> gpurun_out/summary.txt
timeout 300 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench reference exit $?" >> gpurun_out/summary.txt
timeout 300 python tools/bench_vocoder.py --cpu-baseline > gpurun_out/bench_voc.json 2> gpurun_out/bench_voc.err; echo "bench vocoder exit $?" >> gpurun_out/summary.txt
timeout 300 python tools/bench_vocoder.py --batch 1 --seconds 5 > gpurun_out/bench_voc_b1.json 2>> gpurun_out/bench_voc.err; echo "bench vocoder b1 exit $?" >> gpurun_out/summary.txt
timeout 120 python tools/profile_pass.py --passes 1 > gpurun_out/plain.log 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_pass.csv python tools/profile_pass.py --passes 1 > gpurun_out/ncu_list.log 2>&1
echo "launch list exit $?" >> gpurun_out/summary.txt
timeout 120 python tools/profile_vocoder.py > gpurun_out/voc_plain.log 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'voc_conv|k_voc_act|k_voc_post|k_voc_mel_in|k_voc_to_bf16' -c 300 --csv --log-file gpurun_out/launches_vocoder.csv python tools/profile_vocoder.py > gpurun_out/ncu_voc_list.log 2>&1
echo "vocoder launch list exit $?" >> gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_voc_act -s 40 -c 2 -o gpurun_out/prof_voc_act -f python tools/profile_vocoder.py > gpurun_out/ncu_voc3.log 2>&1
echo "ncu act exit $?" >> gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:voc_conv_halo -s 6 -c 4 -o gpurun_out/prof_voc_conv_wide -f python tools/profile_vocoder.py > gpurun_out/ncu_voc1.log 2>&1
echo "ncu conv wide exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -3 gpurun_out/t_all.log; tail -4 gpurun_out/smoke.log
