#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== pytest msda"; timeout 900 python -m pytest tests/test_msda_gpu.py -q -x 2>&1 | tail -15 | tee gpurun_out/pytest_msda.log
echo "=== opbench fwd iters sweep (variant = iters+1; 1 = generic)"
timeout 900 python tools/opbench.py --cases c2_enc_model,c2_enc_uniform,c2_dec,c1_enc_model,c2_enc_model_n2 --variants 1,0,2,3,5,9 --bwd-variants 1,0,2,3,5,9 --out gpurun_out/opbench_r3.json 2>&1 | cut -c1-330 | tee gpurun_out/opbench_r3.log
echo "=== ncu full (auto variants)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:msda_ -c 4 \
   -o gpurun_out/prof_op_r3 -f python tools/opbench.py --once --cases c2_enc_model --variants 0 --bwd-variants 0 > gpurun_out/ncu_full3.log 2>&1
ls -la gpurun_out | tail -5
