#!/usr/bin/env bash
# wgrad N-stacked swap (11=2) + window patch pass (12=1): parity (short timeouts), then bench A/B
set -u
mkdir -p gpurun_out
timeout 500 python -m pytest tests -q -m gpu --timeout 120 -x -k "variants or head or tracker or adam or mhcrnn or multiview" > gpurun_out/r2_tests_c1.log 2>&1
rc=$?; echo "head tests exit $rc"; tail -6 gpurun_out/r2_tests_c1.log
if [ $rc -ne 0 ]; then grep -n "Error\|error\|assert" gpurun_out/r2_tests_c1.log | head -20; fi
for v in "11=2,12=1" "11=1,12=1" "11=2,12=0" "11=1,12=0"; do
  LPB_TUNE="$v" timeout 400 python bench.py --steps 20 --warmup 5 --no-flat --no-cpu-baseline > "gpurun_out/r2_c1_${v//[=,]/_}.json" 2>/dev/null
  python - "$v" <<'PY'
import json,sys,re
v=sys.argv[1]
try:
    d=json.load(open("gpurun_out/r2_c1_%s.json"%re.sub("[=,]","_",v)))
    print(v, " ms/step", round(d["ms_per_step"],4), "bwd_lab", d["stages"]["head_bwd_labeled"]["ms"], "bwd_unl", d["stages"]["head_bwd_unlabeled"]["ms"])
except Exception as e: print(v, " failed", e)
PY
done
