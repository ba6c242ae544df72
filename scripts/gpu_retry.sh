#!/bin/bash
# usage: scripts/gpu_retry.sh <timeout_s> <logfile> '<command>'   -- retries while the pod answers busy (exit 3)
T=$1; LOG=$2; shift 2
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@" > "$LOG" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$LOG"; then exit $rc; fi
  sleep 90
done
exit 3
