#!/bin/bash
# one gpurun --gpus 8 call: the bench contract at N = 8 (weak scaling, one rank per GPU over NCCL)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus8.txt 2>&1
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err; echo "bench 8 gpu exit $?" > gpurun_out/summary8.txt
cat gpurun_out/summary8.txt; tail -c 1500 gpurun_out/bench_8gpu.json; tail -5 gpurun_out/bench_8gpu.err
