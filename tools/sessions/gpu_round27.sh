#!/bin/bash
set -u
timeout 600 python -m pytest tests/test_train_step_gpu.py -q 2>&1 | tail -5
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_n1_r27.json | cut -c1-230
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --torch-adamw 2>/dev/null | tee gpurun_out/bench_n1_r27_torchadamw.json | cut -c1-230
