#!/bin/bash
# First GPU visit: parity tests, smoke, op micro-benchmark, ncu launch list + full capture.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/gpu.txt
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "=== opbench"; timeout 900 python tools/opbench.py --variants 0,1,2 2>&1 | tee gpurun_out/opbench.log | cut -c1-400
echo "=== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:msda_ -c 16 --csv \
   --log-file gpurun_out/launches_op.csv python tools/opbench.py --once --cases c2_enc_model,c2_dec > gpurun_out/ncu_launch.log 2>&1
echo "=== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:msda_ -c 8 \
   -o gpurun_out/prof_op_r1 -f python tools/opbench.py --once --cases c2_enc_model > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
