#!/bin/bash
# ncu --set full of the TMA tile kernel on the C2 encoder problem (one launch)
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:msda_fwd_enc_tma -s 3 -c 1 -f -o gpurun_out/r2_tile_fwd \
  python tools/opbench.py --cases c2_enc_model --variants -1 --bwd-variants 20 --iters 3 --out gpurun_out/tmp_ob.json > gpurun_out/ncu_tile.log 2>&1
tail -5 gpurun_out/ncu_tile.log
ls -la gpurun_out/*.ncu-rep
