#!/usr/bin/env bash
# usage: scripts/gpurun_retry.sh <timeout_s> '<command>'  -- retries while the pod answers "transient" (nothing charged)
T=$1; shift
for i in $(seq 1 60); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1)
  echo "$out" | tail -150
  if ! echo "$out" | grep -q "status=transient"; then exit 0; fi
  echo "[retry $i] transient; sleeping 150 s"
  sleep 150
done
