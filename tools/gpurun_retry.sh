#!/bin/bash
# retry a gpurun call while the pod answers "busy" (exit 3: nothing charged)
# usage: tools/gpurun_retry.sh <timeout_s> '<command>'
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$1" -- "$2"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3
