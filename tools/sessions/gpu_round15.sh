#!/bin/bash
set -u
python - <<'PY'
import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tools')
import opbench
from trackformer_b200 import ext
m = ext.load(); dev = torch.device('cuda:0')
(v, sh, loc, attn, gout), dims, fb, bb = opbench.make_case('c2_enc_model', dev)
m.set_variant(0, 0); a = m.ms_deform_attn_forward(v, sh, loc, attn, 64)
m.set_variant(100, 0); b = m.ms_deform_attn_forward(v, sh, loc, attn, 64)
print('lane-layout max abs diff vs default:', float((a - b).abs().max()))
PY
timeout 600 python tools/opbench.py --cases c2_enc_model,c2_enc_uniform,c2_dec --variants 0,100 --bwd-variants 0 --out gpurun_out/opbench_r15.json 2>&1 | grep '"fwd"' | cut -c1-60,190-300
