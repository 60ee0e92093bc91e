#!/usr/bin/env bash
# one --set full capture of every hot-path kernel of one train step (first step of bench.py)
set -uo pipefail
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:"k1a|k1b|decode_fwd|decode_bwd_window|g2_build|wgrad|b2d|b3a|plane_dot|heatmap_mse" -c 26 \
  -o gpurun_out/prof_r01_step -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-flat > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-200
ls -la gpurun_out/prof_r01_step.ncu-rep
echo "== bench (plain)"; timeout 400 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/bench.log; cut -c1-400 gpurun_out/bench.log
