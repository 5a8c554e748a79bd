#!/bin/bash
# Round 2, call 3: second-generation run forward (variants 120 / 121), fixed MN-major tcgen05 operands, regenerated C2 golden.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "=== tcgen05 GEMM"
timeout 300 python -m pytest tests/test_tf32_gemm_gpu.py -q --maxfail=20 --tb=short > gpurun_out/c3_gemm_tests.log 2>&1; tail -4 gpurun_out/c3_gemm_tests.log
timeout 300 python tools/gemm_bench.py --iters 10 > gpurun_out/c3_gemm_bench.log 2>&1; tail -12 gpurun_out/c3_gemm_bench.log | cut -c1-300
echo "=== op tests"
timeout 1500 python -m pytest tests/test_msda_gpu.py -q --maxfail=30 --tb=short > gpurun_out/c3_op_tests.log 2>&1; tail -40 gpurun_out/c3_op_tests.log | cut -c1-200
echo "=== opbench"
timeout 600 python tools/opbench.py --cases c2_enc_init,c2_enc_model,c2_enc_uniform,c2_enc_init_n2 \
   --variants 20,120,121 --bwd-variants 20 --iters 20 --out gpurun_out/r2_opbench_v2.json > gpurun_out/c3_opbench.log 2>&1
python - <<'PY'
import json
rows=json.load(open('gpurun_out/r2_opbench_v2.json'))
for r in rows: print(r['case'], r['kind'], r['variant'], r['cold_us'], r['warm_us'])
PY
echo "=== model parity (the three that failed)"
timeout 900 python -m pytest tests/test_model_parity_gpu.py -q -k "train_step_c2 or bench_pipeline" --tb=short > gpurun_out/c3_parity.log 2>&1; tail -30 gpurun_out/c3_parity.log | cut -c1-300
