#!/bin/bash
# one replayed training step under ncu (kernel nodes of the step graph + the optimizer launches): launch list at HEAD
set -u
mkdir -p gpurun_out
export TFB200_PROFILE_STEP=1
for cfg in default allnew; do
  if [ $cfg = allnew ]; then export TFB200_FUSED_LOSS=1 TFB200_TCGEN05_LINEAR=1 TFB200_FUSED_PREP=1; fi
  timeout 900 ncu --profile-from-start off --graph-profiling node --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r2_step_launches_$cfg.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_step_$cfg.log 2>&1
  wc -l gpurun_out/r2_step_launches_$cfg.csv
done
