#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:msda_bwd_enc_tma -s 3 -c 1 -f -o gpurun_out/r2_tile_bwd \
  python tools/opbench.py --cases c2_enc_init --variants 20 --bwd-variants -1 --iters 3 --out gpurun_out/tmp_ob.json > gpurun_out/ncu_tile_bwd.log 2>&1
tail -3 gpurun_out/ncu_tile_bwd.log | cut -c1-200
