#!/usr/bin/env bash
# full capture of the window patch kernel
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'g2_patch' -s 2 -c 1 \
  -o gpurun_out/r2_patch python bench.py --profile-step --steps 2 --warmup 1 --no-graph --serial-chains > gpurun_out/r2_patch_ncu.log 2>&1
echo "ncu exit $?"
ncu -i gpurun_out/r2_patch.ncu-rep --page raw --csv > gpurun_out/r2_patch_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_patch.ncu-rep --page details 2>/dev/null | grep -v "^ *$" | head -150
