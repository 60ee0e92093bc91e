#!/usr/bin/env bash
# b3a TMA store: parity (short timeouts: a barrier bug would hang), then bench A/B
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests -q -m gpu --timeout 120 -x -k "variants or head or tracker or adam" > gpurun_out/r2_tests_b8.log 2>&1
rc=$?; echo "head tests exit $rc"; tail -6 gpurun_out/r2_tests_b8.log
if [ $rc -ne 0 ]; then grep -n "Error\|error\|assert" gpurun_out/r2_tests_b8.log | head -20; fi
for v in "11=1" "11=0"; do
  LPB_TUNE="$v" timeout 400 python bench.py --steps 20 --warmup 5 --no-flat --no-cpu-baseline > "gpurun_out/r2_wg_${v//[=,]/_}.json" 2>/dev/null
  python - "$v" <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.load(open("gpurun_out/r2_wg_%s.json"%v.replace("=","_")))
    print(v, " ms/step", round(d["ms_per_step"],4), "bwd_lab", d["stages"]["head_bwd_labeled"]["ms"], "bwd_unl", d["stages"]["head_bwd_unlabeled"]["ms"])
except Exception as e: print(v, " failed", e)
PY
done
