#!/bin/bash
# input prefetch: train-step tests + default bench
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_step_gpu.py -q -m gpu --tb=short > gpurun_out/prefetch_tests.log 2>&1; tail -5 gpurun_out/prefetch_tests.log | cut -c1-300
timeout 900 python bench.py 2>gpurun_out/bench_prefetch.err | tee gpurun_out/bench_prefetch.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['value'],2), 'fps', round(d['ms_per_step'],3), 'ms  e2e', round(d['e2e']['value'],2), round(d['e2e']['ms_per_step'],3), 'var_gt', round(d['variable_gt']['value'],2), d['clocks'])" || tail -5 gpurun_out/bench_prefetch.err
