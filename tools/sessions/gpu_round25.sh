#!/bin/bash
set -u
timeout 600 python -m pytest tests/test_tracker_gpu.py tests/test_fused_norm_gpu.py -q -x 2>&1 | tail -15
timeout 300 python tools/track_bench.py 2>gpurun_out/track_bench_r25.err | tee gpurun_out/track_bench_r25.json | cut -c1-900
tail -3 gpurun_out/track_bench_r25.err
timeout 1200 python -m pytest tests -q -m gpu -x --deselect tests/test_tracker_gpu.py --deselect tests/test_fused_norm_gpu.py 2>&1 | tail -4
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_n1_r25.json | cut -c1-260
