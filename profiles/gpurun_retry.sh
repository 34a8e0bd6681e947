#!/bin/bash
# gpurun with retries while the pod has no free slot (exit 3 / "transient"): usage gpurun_retry.sh <timeout_s> [--gpus N] -- '<command>'
T=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" "$@" 2>&1); rc=$?
  echo "$out" | tail -80
  if echo "$out" | grep -q "status=transient"; then sleep 90; continue; fi
  exit $rc
done
exit 3
