#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== bench graphs v2"; timeout 1200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_err.log | tee gpurun_out/bench_n1_graph2.json | cut -c1-700
grep -v -i warning gpurun_out/bench_err.log | tail -8
echo "=== model parity gpu"; timeout 900 python -m pytest tests/test_model_parity_gpu.py -q 2>&1 | tail -3
echo "=== op shapes"; timeout 900 python tools/op_shapes_profile.py 2>/dev/null > gpurun_out/op_shapes.txt; cut -c1-250 gpurun_out/op_shapes.txt | head -90
