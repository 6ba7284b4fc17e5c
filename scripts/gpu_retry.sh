#!/bin/bash
# Development aid: runs one gpurun call, retrying while the pod has no free GPU slot.
#   bash scripts/gpu_retry.sh <log file> <timeout seconds> '<command>'
log=$1; to=$2; shift 2
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  if grep -q "status=ok\|status=failed\|status=timeout" $log; then break; fi
  sleep 60
done
tail -8 $log | cut -c1-300
