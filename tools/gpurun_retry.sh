#!/bin/bash
# usage: bash tools/gpurun_retry.sh <logfile> [gpurun args...] -- '<command>'
# Retries a gpurun call while the pod answers "busy / draining" (exit code 3, nothing charged).
log=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if ! grep -q "status=transient" "$log"; then exit $rc; fi
  sleep 90
done
