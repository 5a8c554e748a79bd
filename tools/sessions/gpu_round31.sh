#!/bin/bash
set -u
timeout 600 python -m pytest tests/test_msda_gpu.py -q -x 2>&1 | tail -4
timeout 200 python tools/opbench.py --cases c5_enc_model,c5_dec --variants 0,1 --out gpurun_out/opbench_c5_r31.json 2>&1 | grep '"fwd"' | cut -c1-330
