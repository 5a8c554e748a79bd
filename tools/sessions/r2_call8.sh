#!/bin/bash
# Round 2, call 8: TMA-staged encoder tile kernels, forward + backward: parity + timing
set -u
mkdir -p gpurun_out
echo "=== tiled encoder tests"
timeout 900 python -m pytest tests/test_msda_gpu.py -q -k "tiled" --maxfail=30 --tb=short > gpurun_out/c8_tile_tests.log 2>&1; tail -40 gpurun_out/c8_tile_tests.log | cut -c1-200
echo "=== opbench"
timeout 600 python tools/opbench.py --cases c2_enc_init,c2_enc_model,c2_enc_init_n2 \
   --variants 20,-1 --bwd-variants 20,-1,-3 --iters 20 --out gpurun_out/r2_opbench_v4.json > gpurun_out/c8_opbench.log 2>&1
python - <<'PY'
import json
rows=json.load(open('gpurun_out/r2_opbench_v4.json'))
for r in rows: print(r['case'], r['kind'], r['variant'], r['cold_us'], r['warm_us'])
PY
tail -3 gpurun_out/c8_opbench.log | cut -c1-300
