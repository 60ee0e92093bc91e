#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"wgrad|b3a|b2d|relayout|k1a" --launch-skip 5 -c 5 -o gpurun_out/prof_bwd -f python scripts/bwd_only.py > gpurun_out/ncu_bwd.log 2>&1
tail -3 gpurun_out/ncu_bwd.log | cut -c1-200
ls -la gpurun_out/prof_bwd.ncu-rep
