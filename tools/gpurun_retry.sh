#!/bin/bash
# Resubmits a gpurun call while the pod answers "busy" (exit code 3: nothing charged).
# Usage: gpurun_retry.sh <timeout> '<command>' [gpus]
for i in $(seq 1 30); do
  if [ -n "$3" ]; then
    /usr/local/graft/bin/gpurun --gpus "$3" --timeout "$1" -- "$2"
  else
    /usr/local/graft/bin/gpurun --timeout "$1" -- "$2"
  fi
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
