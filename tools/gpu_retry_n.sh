#!/bin/bash
# usage: tools/gpu_retry_n.sh <gpus> <log> <timeout-seconds> <command...>  -- multi-GPU variant of gpu_retry.sh
n=$1; shift; log=$1; shift; to=$1; shift
for i in $(seq 1 60); do
  /usr/local/graft/bin/gpurun --gpus "$n" --timeout "$to" -- "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
