#!/bin/bash
set -u
timeout 600 python -m pytest tests/test_tracker_gpu.py tests/test_fused_norm_gpu.py tests/test_train_step_gpu.py -q 2>&1 | tail -15
timeout 120 python tools/wgrad_bench.py 2>&1 | tail -6
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_n1_r26.json | cut -c1-230
TFB200_WGRAD_SPLITK=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_n1_r26_nosplit.json | cut -c1-230
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --torch-adamw 2>/dev/null | tee gpurun_out/bench_n1_r26_torchadamw.json | cut -c1-230
timeout 300 python tools/opbench.py --cases c5_enc_model,c5_dec --variants 0 --out gpurun_out/opbench_c5_r26.json 2>&1 | tail -8
timeout 400 python tools/track_bench.py --multi-frame --height 1080 --width 1920 --tracks 300 --frames 20 --warmup 5 2>gpurun_out/track_bench_c5.err | tee gpurun_out/track_bench_c5_r26.json | cut -c1-700
tail -3 gpurun_out/track_bench_c5.err
timeout 1200 python -m pytest tests -q -m gpu -x --deselect tests/test_tracker_gpu.py --deselect tests/test_fused_norm_gpu.py --deselect tests/test_train_step_gpu.py 2>&1 | tail -4
