#!/bin/bash
set -u
timeout 900 python -m pytest tests/test_lsa_gpu.py tests/test_train_step_gpu.py tests/test_model_parity_gpu.py -q -x 2>&1 | tail -6
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_n1_r20.json | cut -c1-260
