#!/bin/bash
# one gpurun --gpus 2 call: bench contract + configs sweep at N = 2
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench 2 gpu exit $?" > gpurun_out/summary2.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/sweep.py > gpurun_out/sweep_2gpu.jsonl 2> gpurun_out/sweep2.err; echo "sweep 2 gpu exit $?" >> gpurun_out/summary2.txt
cat gpurun_out/summary2.txt; cut -c1-260 gpurun_out/sweep_2gpu.jsonl; python -c "
import json; d=json.loads(open('gpurun_out/bench_2gpu.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['n_gpus'])"
