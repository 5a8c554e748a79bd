#!/bin/bash
# weighted loss total from the stacked vectors: train-step + bench-pipeline tests, step launch list, bench
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_step_gpu.py tests/test_model_parity_gpu.py -q -m gpu -k "train_step or bench_pipeline or graph or prefetch or flat" --tb=short > gpurun_out/loss_total_tests.log 2>&1; tail -3 gpurun_out/loss_total_tests.log | cut -c1-200
export TFB200_PROFILE_STEP=1
timeout 900 ncu --profile-from-start off --graph-profiling node --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r2_step_launches_final.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_step_final.log 2>&1
wc -l gpurun_out/r2_step_launches_final.csv
unset TFB200_PROFILE_STEP
timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>gpurun_out/bench_loss_total.err | tee gpurun_out/bench_loss_total.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['value'],2), 'fps', round(d['ms_per_step'],3), 'ms  e2e', round(d['e2e']['value'],2), 'var_gt', round(d['variable_gt']['value'],2), d['roofline']['frac'], d['gpu_launches'])"
