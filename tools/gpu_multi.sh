#!/bin/bash
# multi-GPU bench exactly as the driver launches it (fail-fast timeouts: a hang must not hold N GPUs)
set -u
N=${1:-2}
BPG=${2:-1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv | head -9
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps 20 --warmup 3 --batch-per-gpu $BPG 2>gpurun_out/bench_n${N}_b${BPG}_err.log | tee gpurun_out/bench_n${N}_b${BPG}.json | cut -c1-400
grep -v -i "warning\|warn(\|OMP_NUM\|\*\*\*\*\|run_backward" gpurun_out/bench_n${N}_b${BPG}_err.log | tail -6
