#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit code 3: nothing charged).
# usage: bash tools/gpurun_retry.sh <logfile> [gpurun args...] -- '<command>'
log=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$log"; then exit $rc; fi
  sleep 120
done
exit 3
