#!/bin/bash
set -u
timeout 900 python -m pytest tests/test_train_step_gpu.py -q -x 2>&1 | tail -4
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_err.log | tee gpurun_out/bench_n1_r21.json | cut -c1-260
grep -v -i warning gpurun_out/bench_err.log | tail -4
