#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python tools/l1_gather_peak.py
timeout 900 python -m pytest tests/test_train_step_gpu.py -q -x 2>&1 | tail -5
