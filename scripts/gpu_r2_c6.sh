#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 500 python -m pytest tests -q -m gpu --timeout 120 -x -k "variants or head or tracker or adam or mhcrnn or multiview or windows" > gpurun_out/r2_tests_c8.log 2>&1
rc=$?; echo "head tests exit $rc"; tail -6 gpurun_out/r2_tests_c8.log
if [ $rc -ne 0 ]; then grep -n "Error\|error\|assert" gpurun_out/r2_tests_c8.log | head -20; exit 1; fi
for v in "14=1" "14=0"; do
  LPB_TUNE="$v" timeout 400 python bench.py --steps 20 --warmup 5 --no-flat --no-cpu-baseline > "gpurun_out/r2_c8_${v//[=,]/_}.json" 2>/dev/null
  python - "$v" <<'PY'
import json,sys,re
v=sys.argv[1]
try:
    d=json.load(open("gpurun_out/r2_c8_%s.json"%re.sub("[=,]","_",v)))
    print(v, " ms/step", round(d["ms_per_step"],4), "fwd-only", round(d["forward_only"]["ms_per_step"],4), "head_fwd", d["stages"]["head_fwd"]["ms"], "bwd_lab", d["stages"]["head_bwd_labeled"]["ms"], "bwd_unl", d["stages"]["head_bwd_unlabeled"]["ms"])
    f=d.get("fresh_init_regime")
    if f: print("   fresh", round(f["value"]), f["stages"])
except Exception as e: print(v, " failed", e)
PY
done
