#!/bin/bash
# Round 2, call 12: single-launch small column sums, per-step seed pool, decoder linears through fused_linear
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fused_norm_gpu.py tests/test_train_step_gpu.py tests/test_model_parity_gpu.py tests/test_tf32_gemm_gpu.py -q --tb=short > gpurun_out/c12_tests.log 2>&1; tail -8 gpurun_out/c12_tests.log | cut -c1-200
for i in 1 2; do
timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>gpurun_out/bench_c12_$i.err | tee gpurun_out/bench_c12_$i.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['value'],2), 'fps', round(d['ms_per_step'],3), 'ms  e2e', round(d['e2e']['value'],2), 'var_gt', round(d['variable_gt']['value'],2), 'launches', d['gpu_launches'], d['parity_vs_reference_c2'], d['roofline']['frac'], d.get('roofline_dense',{}).get('frac'), d.get('roofline_e2e',{}).get('frac'))" || tail -3 gpurun_out/bench_c12_$i.err
done
