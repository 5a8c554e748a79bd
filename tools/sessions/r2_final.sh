#!/bin/bash
# what the driver runs at round end: GPU suite, smoke(), default bench (both arms); plus the C5 tracking latency
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 2400 python -m pytest tests/ -q -m gpu --tb=short > gpurun_out/final_gpu_suite.log 2>&1; tail -6 gpurun_out/final_gpu_suite.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py 2>gpurun_out/bench_final.err | tee gpurun_out/bench_final.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['value'],2), 'fps', round(d['ms_per_step'],3), 'ms  e2e', round(d['e2e']['value'],2), 'var_gt', round(d['variable_gt']['value'],2), 'cpu', d['cpu_baseline'], d['roofline']['frac'], d['roofline']['kernel'], d['clocks'])" || tail -5 gpurun_out/bench_final.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-300
timeout 600 python tools/track_bench.py --multi-frame --height 1080 --width 1920 --tracks 300 --frames 20 --warmup 5 --skip-eager > gpurun_out/track_bench_c5_device.json 2> gpurun_out/track_bench_c5.err; cut -c1-700 gpurun_out/track_bench_c5_device.json; tail -2 gpurun_out/track_bench_c5.err
