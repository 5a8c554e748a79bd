#!/bin/bash
# Round 2, call 14: small-attention kernels: parity, then the step with / without them
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_small_attn_gpu.py -q --tb=short > gpurun_out/c14_attn_tests.log 2>&1; tail -25 gpurun_out/c14_attn_tests.log | cut -c1-220
timeout 900 python -m pytest tests/test_model_parity_gpu.py tests/test_train_step_gpu.py tests/test_tracker_gpu.py -q --tb=short > gpurun_out/c14_model_tests.log 2>&1; tail -8 gpurun_out/c14_model_tests.log | cut -c1-220
run() {
  local name=$1; shift
  env "$@" timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>gpurun_out/bench_c14_${name}.err | tee gpurun_out/bench_c14_${name}.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$name', round(d['value'],2), 'fps', round(d['ms_per_step'],3), 'ms  e2e', round(d['e2e']['value'],2), 'launches/step', d['gpu_launches']//40)" || tail -3 gpurun_out/bench_c14_${name}.err
}
run attn_on  TFB200_SMALL_ATTN=1
run attn_off TFB200_SMALL_ATTN=0
run attn_on2 TFB200_SMALL_ATTN=1
