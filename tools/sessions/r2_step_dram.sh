#!/bin/bash
# one replayed step under ncu with DRAM byte counters per kernel node (own fused kernels: algorithmic vs moved bytes)
set -u
mkdir -p gpurun_out
export TFB200_PROFILE_STEP=1
timeout 1500 ncu --profile-from-start off --graph-profiling node --clock-control none --csv \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,smsp__inst_executed.sum \
    --log-file gpurun_out/r2_step_dram_head.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_step_dram.log 2>&1
wc -l gpurun_out/r2_step_dram_head.csv
