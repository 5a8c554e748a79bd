#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== pytest fused norm + model parity"; timeout 1200 python -m pytest tests/test_fused_norm_gpu.py tests/test_model_parity_gpu.py -q -x 2>&1 | tail -6
echo "=== bench"; timeout 1200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_err.log | tee gpurun_out/bench_n1_r8.json | cut -c1-330
grep -v -i warning gpurun_out/bench_err.log | tail -5
echo "=== step profile"; timeout 900 python tools/step_profile.py --steps 10 2>/dev/null | tee gpurun_out/step_profile_r8.txt | cut -c1-45,150-240 | head -45
