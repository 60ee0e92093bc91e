#!/usr/bin/env bash
# final captures of the round: tests, smoke, bench line, A/B, configs, reference arm, ncu launch list + full capture
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > gpurun_out/r02_gpu_box.txt 2>&1
timeout 1500 python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/r2_tests.log 2>&1
rc=$?; echo "tests exit $rc"; tail -3 gpurun_out/r2_tests.log
if [ $rc -ne 0 ]; then grep -n "Error\|error\|assert\|FAILED" gpurun_out/r2_tests.log | head -30; exit 1; fi
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1
echo "smoke exit $?"; tail -2 gpurun_out/r2_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_train.json 2> gpurun_out/r02_bench_train.err
echo "bench exit $?"; tail -c 600 gpurun_out/r02_bench_train.json; echo
for v in "serial" "11=1" "12=0" "15=0" "10=0" "7=0"; do
  if [ "$v" = "serial" ]; then fl="--serial-chains"; tn=""; else fl=""; tn="$v"; fi
  LPB_TUNE="$tn" timeout 400 python bench.py --steps 10 --warmup 3 --no-flat --no-cpu-baseline $fl > "gpurun_out/r02_ab_${v//[=,]/_}.json" 2>/dev/null
  python - "$v" <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.load(open("gpurun_out/r02_ab_%s.json"%v.replace("=","_")))
    print("A/B", v, " ms/step", round(d["ms_per_step"],4), {k: s["ms"] for k,s in d["stages"].items()})
except Exception as e: print("A/B", v, " failed", e)
PY
done
timeout 300 python bench.py --profile-step --kineto --steps 5 --warmup 3 --no-graph --serial-chains > gpurun_out/r02_kineto_step.json 2>/dev/null; echo "kineto exit $?"
timeout 600 python scripts/bench_configs.py > gpurun_out/r2_configs.log 2>&1; tail -c 1500 gpurun_out/r2_configs.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null; tail -c 400 gpurun_out/r02_bench_reference_arm.json; echo
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_train_step.csv \
  python bench.py --profile-step --steps 2 --warmup 1 --no-graph --serial-chains > gpurun_out/r02_ncu_list.log 2>&1
echo "ncu launch list exit $?"
timeout 1200 ncu --set full --clock-control none --import-source on \
  -k regex:'k1a_shuffle|convt_rows|decode_fwd|decode_bwd|wgrad_kernel|b3a_dgrad|b2d_dgrad|g2_build|g2_patch|plane_dot|heatmap_mse_from_kp|adam_step|head_prep' -s 31 -c 31 \
  -o gpurun_out/r02_full python bench.py --profile-step --steps 1 --warmup 1 --no-graph --serial-chains > gpurun_out/r02_ncu_full.log 2>&1
echo "ncu full exit $?"
ncu -i gpurun_out/r02_full.ncu-rep --page raw --csv > gpurun_out/r02_full_raw.csv 2>/dev/null
ls -la gpurun_out/ | grep "r02_" | tail -20
