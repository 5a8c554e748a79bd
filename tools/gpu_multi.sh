#!/bin/bash
# multi-GPU bench exactly as the driver launches it
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv | head -9
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps 10 --warmup 3 2>gpurun_out/bench_n${N}_err.log | tee gpurun_out/bench_n${N}.json | cut -c1-600
grep -v -i warning gpurun_out/bench_n${N}_err.log | tail -8
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
   bench.py --impl reference --gpus $N --steps 1 --warmup 0 2>/dev/null | cut -c1-300
