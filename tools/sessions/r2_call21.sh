#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_msda_gpu.py -q -k "fused_prologue or tiled" --tb=short > gpurun_out/c21_tests.log 2>&1; tail -25 gpurun_out/c21_tests.log | cut -c1-220
timeout 900 python -m pytest tests/test_model_parity_gpu.py tests/test_train_step_gpu.py -q --tb=short > gpurun_out/c21_model.log 2>&1; tail -5 gpurun_out/c21_model.log | cut -c1-220
run() {
  local name=$1; shift
  env "$@" timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>gpurun_out/bench_c21_${name}.err | tee gpurun_out/bench_c21_${name}.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$name', round(d['value'],2), 'fps', round(d['ms_per_step'],3), 'ms  e2e', round(d['e2e']['value'],2), [(k['kernel'],k['mean_us']) for k in d['msda_kernels']])" || tail -3 gpurun_out/bench_c21_${name}.err
}
run fused_auto TFB200_X=0
run fused_off  TFB200_TILED_ENC_FUSED=0
run fused_on   TFB200_TILED_ENC_FUSED=1
