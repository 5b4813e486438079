#!/bin/bash
# one gpurun --gpus 8 call: 8-GPU weak-scaling bench + the BASELINE configs sweep at 4 and 8 GPUs (512 utterances sharded)
mkdir -p gpurun_out
S=gpurun_out/summary_8.txt; : > $S
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err; echo "bench 8 GPUs exit $?" >> $S
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 tools/sweep.py > gpurun_out/sweep_8gpu.jsonl 2> gpurun_out/sweep_8gpu.err; echo "sweep 8 GPUs exit $?" >> $S
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 tools/sweep.py > gpurun_out/sweep_4gpu.jsonl 2> gpurun_out/sweep_4gpu.err; echo "sweep 4 GPUs exit $?" >> $S
cat $S; cut -c1-330 gpurun_out/bench_8gpu.json; cut -c1-230 gpurun_out/sweep_8gpu.jsonl; cut -c1-230 gpurun_out/sweep_4gpu.jsonl
