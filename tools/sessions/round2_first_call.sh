#!/bin/bash
# First gpurun call of round 2 on this branch: validate the three prepared changes, then A/B each switch.
#   gpurun --timeout 1500 -- 'bash tools/sessions/round2_first_call.sh'
set -u
timeout 600 python -m pytest tests/test_fused_bn_gpu.py tests/test_fused_norm_gpu.py tests/test_train_step_gpu.py tests/test_model_parity_gpu.py -q 2>&1 | tail -6
run() {  # name, env...
  local name=$1; shift
  env "$@" timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_r2_${name}.json | cut -c1-170
}
run all_on      TFB200_FUSED_BN=1 TFB200_GATHER_GRADS=1 TFB200_LN_SEEDED=1
run no_fused_bn TFB200_FUSED_BN=0 TFB200_GATHER_GRADS=1 TFB200_LN_SEEDED=1
run no_gather   TFB200_FUSED_BN=1 TFB200_GATHER_GRADS=0 TFB200_LN_SEEDED=1
run no_seeded   TFB200_FUSED_BN=1 TFB200_GATHER_GRADS=1 TFB200_LN_SEEDED=0
run all_off     TFB200_FUSED_BN=0 TFB200_GATHER_GRADS=0 TFB200_LN_SEEDED=0
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
