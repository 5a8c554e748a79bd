#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fused_bn_gpu.py tests/test_fused_norm_gpu.py tests/test_model_parity_gpu.py tests/test_train_step_gpu.py -q --tb=short > gpurun_out/c20_tests.log 2>&1; tail -6 gpurun_out/c20_tests.log | cut -c1-200
for i in 1 2; do
timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>gpurun_out/bench_c20_$i.err | tee gpurun_out/bench_c20_$i.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['value'],2), 'fps', round(d['ms_per_step'],3), 'ms  e2e', round(d['e2e']['value'],2), 'var_gt', round(d['variable_gt']['value'],2), [(k['kernel'],k['mean_us']) for k in d['msda_kernels']])" || tail -3 gpurun_out/bench_c20_$i.err
done
