#!/bin/bash
# retries a gpurun call while the pod answers "busy" (status=transient, nothing charged); usage: gpurun_retry.sh <timeout> <command>
for i in $(seq 1 12); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$1" -- "$2" 2>&1)
  if ! echo "$out" | grep -q "status=transient"; then echo "$out"; exit 0; fi
  sleep 60
done
echo "$out"
