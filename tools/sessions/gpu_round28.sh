#!/bin/bash
set -u
timeout 300 python tools/optimizer_bench.py 2>&1 | tail -2
timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_n1_r28a.json | cut -c1-200
timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --torch-adamw 2>/dev/null | tee gpurun_out/bench_n1_r28b.json | cut -c1-200
timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_n1_r28c.json | cut -c1-200
