#!/bin/bash
# usage: scripts/gpurun_retry.sh <tries> <gpurun args...>   -- retries while gpurun answers "busy" (nothing charged)
tries=$1; shift
for i in $(seq 1 $tries); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient"; then
    echo "[retry $i] busy; sleeping 90 s"; sleep 90; continue
  fi
  echo "$out"; exit $rc
done
echo "gave up after $tries tries"; exit 3
