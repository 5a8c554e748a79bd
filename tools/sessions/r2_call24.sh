#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== D = 36 (multi-frame geometry) kernel families"
timeout 600 python tools/opbench.py --cases c5_enc_model,c5_dec --variants 0,100,101 --bwd-variants 0,100,101 --iters 10 --out gpurun_out/r2_opbench_c5.json > gpurun_out/c24_opbench.log 2>&1
python - <<'PY'
import json
rows=json.load(open('gpurun_out/r2_opbench_c5.json'))
for r in rows: print(r['case'], r['kind'], r['variant'], r['cold_us'], r['warm_us'])
PY
tail -2 gpurun_out/c24_opbench.log | cut -c1-200
echo "=== batch 2 per GPU (C4 per-GPU batch)"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --batch-per-gpu 2 2>gpurun_out/bench_c24_b2.err | tee gpurun_out/bench_c24_b2.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['value'],2), 'fps', round(d['ms_per_step'],3), 'ms  e2e', round(d['e2e']['value'],2), [(k['kernel'],k['mean_us']) for k in d['msda_kernels']])" || tail -5 gpurun_out/bench_c24_b2.err
