#!/bin/bash
# Round 2, call 1: validate the merged prep work (fused NHWC FrozenBN, gradient gather, seeded LN dropout),
# A/B each switch in ONE call, then a fresh ncu launch list of the step graph at HEAD.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 900 python -m pytest tests/test_fused_bn_gpu.py tests/test_fused_norm_gpu.py tests/test_train_step_gpu.py tests/test_model_parity_gpu.py -q 2>&1 | tail -8
run() {
  local name=$1; shift
  env "$@" timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>gpurun_out/bench_r2_${name}.err | tee gpurun_out/bench_r2_${name}.json | cut -c1-170
}
run all_on      TFB200_FUSED_BN=1 TFB200_GATHER_GRADS=1 TFB200_LN_SEEDED=1
run no_fused_bn TFB200_FUSED_BN=0 TFB200_GATHER_GRADS=1 TFB200_LN_SEEDED=1
run no_gather   TFB200_FUSED_BN=1 TFB200_GATHER_GRADS=0 TFB200_LN_SEEDED=1
run no_seeded   TFB200_FUSED_BN=1 TFB200_GATHER_GRADS=1 TFB200_LN_SEEDED=0
echo "=== ncu launch list of the step graph (kernel nodes)"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -c 14000 --csv \
  --log-file gpurun_out/r2_launches_head.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
wc -l gpurun_out/r2_launches_head.csv
gzip -f gpurun_out/r2_launches_head.csv
