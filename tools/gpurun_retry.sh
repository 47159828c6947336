#!/bin/bash
# retries a gpurun call while the pod answers "busy" (status=transient, nothing charged)
# usage: gpurun_retry.sh <timeout> <command> [extra gpurun flags, e.g. --gpus 4]
t=$1; cmd=$2; shift 2
for i in $(seq 1 12); do
  out=$(/usr/local/graft/bin/gpurun "$@" --timeout "$t" -- "$cmd" 2>&1)
  if ! echo "$out" | grep -q "status=transient"; then echo "$out"; exit 0; fi
  sleep 60
done
echo "$out"
