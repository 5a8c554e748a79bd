#!/bin/bash
# Round 2, call 10: full GPU suite at HEAD + A/B of the step-level switches (one call, same box)
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "=== full GPU suite"
timeout 2400 python -m pytest tests/ -q -m gpu --tb=short -x --maxfail=15 > gpurun_out/c10_gpu_suite.log 2>&1; tail -25 gpurun_out/c10_gpu_suite.log | cut -c1-200
run() {
  local name=$1; shift
  env "$@" timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>gpurun_out/bench_c10_${name}.err | tee gpurun_out/bench_c10_${name}.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$name', round(d['value'],2), 'fps', round(d['ms_per_step'],3), 'ms  e2e', round(d['e2e']['value'],2), ' msda', d.get('msda_ms_per_step'), [ (k['kernel'],k['mean_us']) for k in d.get('msda_kernels',[])])" || tail -3 gpurun_out/bench_c10_${name}.err
}
run default     TFB200_X=0
run notile      TFB200_TILED_ENC=0
run tilebwd     TFB200_TILED_ENC_BWD=1
run fused_loss  TFB200_FUSED_LOSS=1
run tc_w        TFB200_TCGEN05_LINEAR=1 TFB200_TCGEN05_PARTS=w
run tc_dw       TFB200_TCGEN05_LINEAR=1 TFB200_TCGEN05_PARTS=dw
run tc_fdw      TFB200_TCGEN05_LINEAR=1 TFB200_TCGEN05_PARTS=fdw
run mha_unfused TFB200_MHA_NEED_WEIGHTS=1
run best        TFB200_FUSED_LOSS=1 TFB200_TCGEN05_LINEAR=1 TFB200_TCGEN05_PARTS=w
