#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== pytest msda"; timeout 900 python -m pytest tests/test_msda_gpu.py -q -x 2>&1 | tail -4 | tee gpurun_out/pytest_msda.log
echo "=== step profile (graphs)"; timeout 900 python tools/step_profile.py --steps 10 2>/dev/null | tee gpurun_out/step_profile_graph.txt | cut -c1-200 | head -70
echo "=== step profile batch 2"; timeout 900 python tools/step_profile.py --steps 6 --batch 2 2>/dev/null | grep -E "PHASES|PIPELINED" | tee gpurun_out/step_profile_b2.txt
