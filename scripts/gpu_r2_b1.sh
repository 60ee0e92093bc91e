#!/usr/bin/env bash
# tests + smoke (bounded); variants A/B'd when something fails
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_gpu.txt 2>&1
timeout 1500 python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/r2_tests.log 2>&1
rc=$?
echo "tests exit $rc" >> gpurun_out/r2_tests.log
tail -40 gpurun_out/r2_tests.log
if [ $rc -ne 0 ]; then
  LPB_TUNE="0=0,1=0,2=0,3=0" timeout 1200 python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/r2_tests_oldvariants.log 2>&1
  echo "old-variant tests exit $?" >> gpurun_out/r2_tests_oldvariants.log
  tail -30 gpurun_out/r2_tests_oldvariants.log
fi
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/r2_smoke.log
tail -3 gpurun_out/r2_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-flat --no-cpu-baseline > gpurun_out/r2_bench_quick.json 2> gpurun_out/r2_bench_quick.err
echo "bench exit $?"; tail -c 2500 gpurun_out/r2_bench_quick.json; tail -3 gpurun_out/r2_bench_quick.err
