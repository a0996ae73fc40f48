#!/bin/bash
# gpurun with retries while the pod answers "transient / busy" (nothing is charged for those)
# usage: tools/gpurun_retry.sh <logfile> <gpurun args...>
LOG=$1; shift
for attempt in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  rc=$?
  if grep -q "status=transient" "$LOG" || [ $rc -eq 3 ]; then sleep 100; continue; fi
  exit $rc
done
exit 3
