#!/bin/bash
# Second GPU visit: full parity suite, smoke, first bench line, launch list of the bench step.
set -u
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -30 | tee gpurun_out/pytest_gpu.log
echo "=== pytest model parity (verbose deviations)"; timeout 900 python -m pytest tests/test_model_parity_gpu.py -q -s 2>&1 | grep -E "^\{|tf32|passed|failed|Error" | tee gpurun_out/model_parity.log
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/smoke.log
echo "=== bench N=1"; timeout 1200 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench_err.log | tee gpurun_out/bench_n1.json | cut -c1-1500
tail -5 gpurun_out/bench_err.log
echo "=== bench N=1 strict fp32"; timeout 600 python bench.py --steps 5 --warmup 3 --no-tf32 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_n1_fp32.json | cut -c1-400
echo "=== ncu launch list of the bench step"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv \
   --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-300
ls -la gpurun_out
