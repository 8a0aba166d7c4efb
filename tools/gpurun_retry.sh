#!/bin/bash
# gpurun with retries while the pod is busy (exit code 3 / "transient"): tools/gpurun_retry.sh LOG TIMEOUT CMD...
log=$1; to=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  if ! grep -q "status=transient" $log; then exit 0; fi
  sleep 90
done
