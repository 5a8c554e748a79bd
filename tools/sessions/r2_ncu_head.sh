#!/bin/bash
# ncu --set full at HEAD: TMA tile forward / backward on the C2 encoder call (init-like and trained-like sampling locations),
# and the tracker-step kernel (one launch each); tile-kernel parity tests and the op benchmark of the same build
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_msda_gpu.py -q -m gpu -k "tile or tiled or enc" --tb=short > gpurun_out/head_tile_tests.log 2>&1; tail -3 gpurun_out/head_tile_tests.log | cut -c1-200
for case in c2_enc_init c2_enc_model; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:msda_fwd_enc_tma -s 3 -c 1 -f -o gpurun_out/r2_head_tile_fwd_$case \
    python tools/opbench.py --cases $case --variants -1 --bwd-variants 20 --iters 3 --out gpurun_out/tmp_ob.json > gpurun_out/ncu_head_fwd.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:msda_bwd_enc_tma -s 3 -c 1 -f -o gpurun_out/r2_head_tile_bwd_$case \
    python tools/opbench.py --cases $case --variants 20 --bwd-variants -1 --iters 3 --out gpurun_out/tmp_ob.json > gpurun_out/ncu_head_bwd.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:track_step_kernel -s 2 -c 1 -f -o gpurun_out/r2_head_track_step \
  python -m pytest tests/test_tracker_gpu.py -q -m gpu -k "device_tracker_many" > gpurun_out/ncu_head_track.log 2>&1
tail -2 gpurun_out/ncu_head_track.log | cut -c1-200
timeout 600 python tools/opbench.py --cases c2_enc_model,c2_enc_init --variants 20,-1 --bwd-variants 20,-1 --iters 30 --out gpurun_out/opbench_head.json > gpurun_out/opbench_head.log 2>&1
grep '"kind"' gpurun_out/opbench_head.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['case'], d['kind'], d['variant'], d['cold_us'], d['warm_us'])"
ls -la gpurun_out/*.ncu-rep | awk '{print \$5, \$9}'
