#!/bin/bash
set -u
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
timeout 900 python bench.py 2>/dev/null | tee gpurun_out/bench_n1_r29.json | cut -c1-400
