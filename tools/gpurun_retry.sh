#!/bin/bash
# gpurun with retries while the pod answers "busy / draining" (exit code 3 or status=transient): nothing is charged for those.
# usage: tools/gpurun_retry.sh <logfile> <gpurun args...>
LOG=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  rc=$?
  if grep -q "status=transient" "$LOG" || [ $rc -eq 3 ]; then sleep 90; continue; fi
  exit $rc
done
exit 3
