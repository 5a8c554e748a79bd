#!/bin/bash
# what the driver runs at round end: GPU suite, smoke(), default bench (both arms)
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 1200 python -m pytest tests/ -q -m gpu --tb=short > gpurun_out/final_gpu_suite.log 2>&1; tail -4 gpurun_out/final_gpu_suite.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py 2>gpurun_out/bench_final.err | tee gpurun_out/bench_final.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['value'],2), 'fps', round(d['ms_per_step'],3), 'ms  e2e', round(d['e2e']['value'],2), 'var_gt', round(d['variable_gt']['value'],2), 'cpu', d['cpu_baseline']['value'], d['roofline']['frac'], d['roofline']['kernel'], d['clocks'])" || tail -5 gpurun_out/bench_final.err
