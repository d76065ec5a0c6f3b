#!/bin/bash
# gpurun with patience: retry while the pod answers "busy" (exit 3). Usage: tools/gpurun_retry.sh <tries> <gpurun args...>
tries=$1; shift
for i in $(seq 1 $tries); do
  /usr/local/graft/bin/gpurun "$@"; rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 75
done
exit 3
