#!/bin/bash
# gpurun with patience: retries while the pod answers "busy / transient" (exit code 3, nothing charged).
#   scripts/gpurun_retry.sh <log> [gpurun options] -- '<command>'
log=$1; shift
for attempt in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$log"; then exit $rc; fi
  sleep 90
done
exit 3
