#!/bin/bash
# DeviceTracker on the GPU: parity tests, racecheck / memcheck of the one-CTA kernel, tracking latency host vs device decisions
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tracker_gpu.py -q -m gpu --tb=short > gpurun_out/dt_tests.log 2>&1; tail -8 gpurun_out/dt_tests.log | cut -c1-220
for tool in racecheck memcheck; do
  timeout 600 /usr/local/cuda/bin/compute-sanitizer --tool $tool --kernel-regex kns=track_step_kernel \
    python -m pytest tests/test_tracker_gpu.py -q -m gpu -k "device_tracker_matches_reference or device_tracker_many" --tb=short \
    > gpurun_out/dt_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/dt_$tool.log | tail -4
done
timeout 600 python tools/track_bench.py --skip-eager > gpurun_out/track_bench_c3_device.json 2> gpurun_out/track_bench.err; cut -c1-900 gpurun_out/track_bench_c3_device.json; tail -3 gpurun_out/track_bench.err
