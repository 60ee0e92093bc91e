#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
for c in 0 1 2 4 5 6 3; do
  ST_ONLY=$c timeout 60 python scripts/test_umma_selftest.py 2>&1 | tail -2 | tee -a gpurun_out/selftest.log
  echo "rc=$? (case $c)" | tee -a gpurun_out/selftest.log
done
