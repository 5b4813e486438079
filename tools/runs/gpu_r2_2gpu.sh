#!/bin/bash
# one gpurun --gpus 2 call: second-device handle in one process (per-device kernel attributes) + 2-GPU weak-scaling bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_solve.py -q -m gpu -x -s -k "second_device" > gpurun_out/t_2dev.log 2>&1; echo "pytest second device exit $?" > gpurun_out/summary.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench 2 GPUs exit $?" >> gpurun_out/summary.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/sweep.py > gpurun_out/sweep_2gpu.jsonl 2> gpurun_out/sweep_2gpu.err; echo "sweep 2 GPUs exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -3 gpurun_out/t_2dev.log; cut -c1-300 gpurun_out/bench_2gpu.json; cut -c1-200 gpurun_out/sweep_2gpu.jsonl
