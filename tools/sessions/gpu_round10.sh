#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== pytest msda"; timeout 900 python -m pytest tests/test_msda_gpu.py -q -x 2>&1 | tail -6
echo "=== opbench tiled (-1) vs d32 (0)"; timeout 900 python tools/opbench.py --cases c2_enc_model,c2_enc_uniform,c1_enc_model,c2_enc_model_n2 --variants=-1,0 --bwd-variants 0 --out gpurun_out/opbench_r10.json 2>&1 | cut -c1-200
echo "=== ncu tiled"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile -c 2 -o gpurun_out/prof_tile_r10 -f python tools/opbench.py --once --cases c2_enc_model --variants=-1 --bwd-variants 0 > gpurun_out/ncu_tile.log 2>&1; tail -2 gpurun_out/ncu_tile.log
