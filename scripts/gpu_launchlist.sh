#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_train.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-flat --no-graph > gpurun_out/ncu_list_train.log 2>&1
tail -1 gpurun_out/ncu_list_train.log | cut -c1-200
