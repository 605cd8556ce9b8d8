#!/bin/bash
# usage: gpu_retry.sh <outfile> <gpurun args...>  -- retries while the pod answers busy (exit 3)
out=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$out" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
