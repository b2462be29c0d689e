#!/bin/bash
# gpurun with retries while the pod's GPU slots are busy (exit code 3 = nothing charged)
# usage: scripts/gpurun_retry.sh <timeout_s> <command...>
t=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $t -- "$@" > /tmp/gpurun_last.log 2>&1
  rc=$?
  if grep -q "status=transient" /tmp/gpurun_last.log; then sleep 60; continue; fi
  break
done
tail -60 /tmp/gpurun_last.log
exit $rc
