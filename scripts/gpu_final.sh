#!/usr/bin/env bash
# round-end validation: tests, smoke, bench (both arms), launch list, full ncu capture of one train step
set -uo pipefail
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu_box.txt 2>&1
echo "== pytest gpu"; timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee gpurun_out/smoke.log
echo "== bench"; timeout 500 python bench.py --steps 10 --warmup 3 2> gpurun_out/bench.err | tail -1 > gpurun_out/bench.log; cut -c1-250 gpurun_out/bench.log; tail -2 gpurun_out/bench.err
if [[ "${QUICK:-0}" == "1" ]]; then exit 0; fi
echo "== bench --impl reference"; timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_ref.log; cut -c1-250 gpurun_out/bench_ref.log
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_train.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-flat --no-graph > gpurun_out/ncu_list_train.log 2>&1
tail -1 gpurun_out/ncu_list_train.log | cut -c1-160
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:"k1a|k1b|decode_fwd|decode_bwd_window|g2_build|wgrad|b2d|b3a|plane_dot|heatmap_mse" -c 30 \
  -o gpurun_out/prof_r01_step -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-flat --no-graph > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log | cut -c1-160
ls -la gpurun_out/*.ncu-rep
