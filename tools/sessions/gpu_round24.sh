#!/bin/bash
set -u
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_n1_r24.json | cut -c1-260
