#!/usr/bin/env bash
# round 2, call A: the whole GPU suite (bounded), smoke, bench; kernel variants A/B'd when something fails
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_gpu.txt 2>&1
timeout 2400 python -m pytest tests -q -m gpu --timeout 300 -x > gpurun_out/r2a_tests.log 2>&1
rc=$?
echo "tests exit $rc" >> gpurun_out/r2a_tests.log
tail -30 gpurun_out/r2a_tests.log
if [ $rc -ne 0 ]; then
  # which failures belong to the new kernel variants?  same suite with the round-1 variants
  LPB_TUNE="0=0,1=0,2=0,3=0" timeout 2400 python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/r2a_tests_oldvariants.log 2>&1
  echo "old-variant tests exit $?" >> gpurun_out/r2a_tests_oldvariants.log
  tail -30 gpurun_out/r2a_tests_oldvariants.log
  timeout 2400 python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/r2a_tests_all.log 2>&1
  echo "all (no -x) tests exit $?" >> gpurun_out/r2a_tests_all.log
  tail -40 gpurun_out/r2a_tests_all.log
fi
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/r2a_smoke.log
tail -3 gpurun_out/r2a_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench exit $?"
tail -c 3500 gpurun_out/r2a_bench.json
tail -5 gpurun_out/r2a_bench.err
timeout 600 python scripts/bench_configs.py > gpurun_out/r2a_configs.log 2>&1; tail -c 2500 gpurun_out/r2a_configs.log
timeout 600 python scripts/kp_error_hist.py > gpurun_out/r2a_kp_hist.log 2>&1; tail -c 1500 gpurun_out/r2a_kp_hist.log
LPB_TUNE="0=0,1=0,2=0,3=0" timeout 600 python bench.py --steps 10 --warmup 3 --no-flat --no-cpu-baseline > gpurun_out/r2a_bench_oldvariants.json 2> gpurun_out/r2a_bench_oldvariants.err
echo "bench (round-1 variants) exit $?"
tail -c 1500 gpurun_out/r2a_bench_oldvariants.json
