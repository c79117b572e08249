#!/bin/bash
# tools/ab/gpu_retry.sh <timeout> <command...>: retry gpurun while the pod answers busy (exit 3 / transient)
to=$1; shift
for k in 1 2 3 4 5 6 7 8 9 10 11 12; do
  out=$(/usr/local/graft/bin/gpurun --timeout $to -- "$@" 2>&1)
  echo "$out"
  if ! echo "$out" | grep -q "status=transient"; then exit 0; fi
  sleep 90
done
