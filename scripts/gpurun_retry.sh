#!/bin/bash
# gpurun with retries while the pod has no free slot (exit code 3 = nothing charged).
# usage: scripts/gpurun_retry.sh <log> [gpurun options] -- <command>
log=$1; shift
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 60
done
exit 3
