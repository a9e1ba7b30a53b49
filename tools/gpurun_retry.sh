#!/bin/bash
# dev helper: retry a gpurun call while the pod answers "busy / draining" (exit code 3, nothing charged)
# usage: tools/gpurun_retry.sh <timeout_s> '<command>'
T=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
