#!/bin/bash
# parity tests of the tile kernels + op timings (init-like / trained-like / uniform C2 encoder calls)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_msda_gpu.py -q -m gpu -k "tile or tiled or enc" --tb=short > gpurun_out/head_tile_tests.log 2>&1; tail -3 gpurun_out/head_tile_tests.log | cut -c1-200
timeout 600 python tools/opbench.py --cases c2_enc_init,c2_enc_model,c2_enc_uniform --variants 20,-1 --bwd-variants 20,-1 --iters 30 --out gpurun_out/opbench_head.json > gpurun_out/opbench_head.log 2>&1
grep '"kind"' gpurun_out/opbench_head.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['case'], d['kind'], d['variant'], d['cold_us'], d['warm_us'])"
