#!/bin/bash
set -u
timeout 600 python tools/opbench.py --cases c2_enc_model,c2_dec --variants 0 --bwd-variants 0,10,20,30,40 --iters 20 --out gpurun_out/opbench_r19.json 2>&1 | grep '"bwd"' | python -c "
import sys, json
for l in sys.stdin:
    r=json.loads(l); print(f\"{r['case']:14s} v{r['variant']:3d} cold {r['cold_us']:7.1f} warm {r['warm_us']:7.1f}\")
"
