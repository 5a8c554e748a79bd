#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== pytest"; timeout 1200 python -m pytest tests/test_fused_norm_gpu.py tests/test_model_parity_gpu.py -q -x 2>&1 | tail -4
echo "=== bench"; timeout 1200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_err.log | tee gpurun_out/bench_n1_r11.json | cut -c1-330
grep -v -i warning gpurun_out/bench_err.log | tail -5
echo "=== bench NCHW"; TFB200_CHANNELS_LAST=0 timeout 1200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-250
