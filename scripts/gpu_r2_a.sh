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
# ---- profiles: launch list of one eager train step + full capture of the hot kernels of the second step ----
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_train_step.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-flat --no-graph > gpurun_out/r02_ncu_list.log 2>&1
echo "ncu launch list exit $?"; tail -2 gpurun_out/r02_ncu_list.log
timeout 1200 ncu --set full --clock-control none --import-source on \
  -k regex:'k1a_shuffle|convt_rows|decode_fwd_ring|decode_fwd_warp|wgrad_kernel|b3a_dgrad|b2d_dgrad|g2_build|heatmap_mse_from_kp' -s 20 -c 20 \
  -o gpurun_out/r02_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-flat --no-graph > gpurun_out/r02_ncu_full.log 2>&1
echo "ncu full exit $?"; tail -2 gpurun_out/r02_ncu_full.log
ncu -i gpurun_out/r02_full.ncu-rep --page raw --csv > gpurun_out/r02_full_raw.csv 2>/dev/null
ls -la gpurun_out/ | tail -20
