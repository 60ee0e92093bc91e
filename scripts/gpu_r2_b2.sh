#!/usr/bin/env bash
# measurements: full bench line, variant A/B, other configs, bf16 keypoint error, ncu captures
set -u
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
echo "bench exit $?"; tail -c 3500 gpurun_out/r2_bench.json; tail -3 gpurun_out/r2_bench.err
for v in "0=1" "1=0" "2=0" "3=1" "4=1"; do
  LPB_TUNE="$v" timeout 400 python bench.py --steps 10 --warmup 3 --no-flat --no-cpu-baseline > "gpurun_out/r2_bench_tune_${v//[=,]/_}.json" 2>/dev/null
  echo "tune $v:"; python - "$v" <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.load(open("gpurun_out/r2_bench_tune_%s.json"%v.replace("=","_").replace(",","_")))
    print(" ms/step", round(d["ms_per_step"],4), {k: s["ms"] for k,s in d["stages"].items()})
except Exception as e: print(" failed", e)
PY
done
timeout 600 python scripts/bench_configs.py > gpurun_out/r2_configs.log 2>&1; tail -c 3000 gpurun_out/r2_configs.log
timeout 600 python scripts/kp_error_hist.py > gpurun_out/r2_kp_hist.log 2>&1; tail -c 1500 gpurun_out/r2_kp_hist.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference.json 2>/dev/null; tail -c 800 gpurun_out/r2_bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_train_step.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-flat --no-graph > gpurun_out/r02_ncu_list.log 2>&1
echo "ncu launch list exit $?"
timeout 1200 ncu --set full --clock-control none --import-source on \
  -k regex:'k1a_shuffle|convt_rows|decode_fwd_ring|decode_fwd_warp|wgrad_kernel|b3a_dgrad|b2d_dgrad|g2_build|heatmap_mse_from_kp' -s 20 -c 20 \
  -o gpurun_out/r02_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-flat --no-graph > gpurun_out/r02_ncu_full.log 2>&1
echo "ncu full exit $?"
ncu -i gpurun_out/r02_full.ncu-rep --page raw --csv > gpurun_out/r02_full_raw.csv 2>/dev/null
ls -la gpurun_out/ | tail -25
