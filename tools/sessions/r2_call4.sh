#!/bin/bash
# Round 2, call 4: TMA-staged encoder tile kernel (forward): parity + timing; gemm autograd test with error codes.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "=== tiled encoder tests"
timeout 600 python -m pytest tests/test_msda_gpu.py -q -k "tiled or families_match" --maxfail=30 --tb=short > gpurun_out/c4_tile_tests.log 2>&1; tail -30 gpurun_out/c4_tile_tests.log | cut -c1-220
echo "=== opbench"
timeout 600 python tools/opbench.py --cases c2_enc_init,c2_enc_model,c2_enc_uniform,c2_enc_init_n2,c1_enc_model \
   --variants 20,-1,-2 --bwd-variants 20 --iters 20 --out gpurun_out/r2_opbench_v3.json > gpurun_out/c4_opbench.log 2>&1
python - <<'PY'
import json
rows=json.load(open('gpurun_out/r2_opbench_v3.json'))
for r in rows: print(r['case'], r['kind'], r['variant'], r['cold_us'], r['warm_us'])
PY
tail -3 gpurun_out/c4_opbench.log | cut -c1-300
echo "=== tcgen05 GEMM autograd"
timeout 300 python -m pytest tests/test_tf32_gemm_gpu.py -q --maxfail=20 --tb=short > gpurun_out/c4_gemm_tests.log 2>&1; tail -6 gpurun_out/c4_gemm_tests.log | cut -c1-200
