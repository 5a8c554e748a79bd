#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== step profile"; timeout 900 python tools/step_profile.py --steps 10 2>/dev/null > gpurun_out/step_profile_r16.txt; cut -c1-60,150-235 gpurun_out/step_profile_r16.txt | head -60
echo "=== ncu launch list of bench (graph kernel nodes)"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -s 20000 -c 12000 --csv \
   --log-file gpurun_out/launches_bench_r16.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_r16.log 2>&1
tail -2 gpurun_out/ncu_bench_r16.log | cut -c1-200; wc -l gpurun_out/launches_bench_r16.csv
