#!/bin/bash
# usage: tools/gpu_retry8.sh <gpus> <timeout_s> '<command>'   - retries gpurun --gpus N while the pod answers busy (exit code 3)
N=$1; T=$2; shift; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --gpus "$N" --timeout "$T" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
