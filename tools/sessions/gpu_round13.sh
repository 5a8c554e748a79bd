#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== pytest refcuda + full gpu suite"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -5
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
echo "=== opbench vs reference CUDA kernels"; timeout 900 python tools/opbench.py --cases c2_enc_model,c2_dec,c1_enc_model --variants 0 --bwd-variants 0 --ref --out gpurun_out/opbench_r13.json 2>&1 | cut -c1-60,200-330
echo "=== bench"; timeout 1200 python bench.py --steps 20 --warmup 5 2>gpurun_out/bench_err.log | tee gpurun_out/bench_n1_r13.json | cut -c1-260
