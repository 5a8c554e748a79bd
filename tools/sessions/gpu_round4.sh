#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== pytest msda"; timeout 900 python -m pytest tests/test_msda_gpu.py -q -x 2>&1 | tail -5 | tee gpurun_out/pytest_msda.log
echo "=== diag"; timeout 600 python tools/opbench.py --cases diag_l0_only,diag_l1_only,diag_l2_only,diag_l3_only,diag_l0_local --variants 2 --bwd-variants 2 --out gpurun_out/opbench_diag.json 2>&1 | cut -c1-260
echo "=== bench graphs"; timeout 1200 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench_err.log | tee gpurun_out/bench_n1_graph.json | cut -c1-1200
tail -15 gpurun_out/bench_err.log
