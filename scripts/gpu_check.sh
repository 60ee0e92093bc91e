#!/usr/bin/env bash
# One gpurun call: GPU parity tests + smoke + a short bench. Logs land in gpurun_out/.
set -uo pipefail
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/gpu.txt
echo "== pytest gpu" ; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -15 | tee gpurun_out/smoke.log
echo "== bench" ; timeout 900 python bench.py --steps 5 --warmup 3 ${BENCH_ARGS:-} 2>&1 | tail -5 | tee gpurun_out/bench.log
