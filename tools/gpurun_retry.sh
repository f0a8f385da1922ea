#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-seconds> <command...>; retries while the pod answers busy (exit 3 / transient)
t=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout "$t" -- "$@" > /tmp/gpurun_retry.log 2>&1
  rc=$?
  if grep -q "status=transient" /tmp/gpurun_retry.log || [ $rc -eq 3 ]; then sleep 120; continue; fi
  break
done
tail -15 /tmp/gpurun_retry.log
