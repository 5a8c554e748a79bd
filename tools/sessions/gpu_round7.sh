#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== pytest msda"; timeout 900 python -m pytest tests/test_msda_gpu.py -q -x 2>&1 | tail -3
echo "=== opbench"; timeout 900 python tools/opbench.py --cases c2_enc_model,c2_enc_uniform,c2_dec,c2_enc_model_n2 --variants 0,3 --bwd-variants 0,3 --out gpurun_out/opbench_r7.json 2>&1 | cut -c1-300
