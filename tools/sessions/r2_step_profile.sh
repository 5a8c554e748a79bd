#!/bin/bash
# one replayed training step under ncu (kernel nodes of the step graph + the optimizer launches): launch list at HEAD
set -u
mkdir -p gpurun_out
export TFB200_PROFILE_STEP=1
timeout 900 ncu --profile-from-start off --graph-profiling node --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r2_step_launches_head.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_step_head.log 2>&1
wc -l gpurun_out/r2_step_launches_head.csv
unset TFB200_PROFILE_STEP
timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>gpurun_out/bench_head.err | tee gpurun_out/bench_head.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['value'],2), 'fps', round(d['ms_per_step'],3), 'ms  e2e', round(d['e2e']['value'],2), 'var_gt', round(d['variable_gt']['value'],2), [(k['kernel'],k['mean_us'],k['frac']) for k in d['msda_kernels']], d['roofline']['frac'])"
