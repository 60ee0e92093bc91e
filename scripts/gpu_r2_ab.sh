#!/usr/bin/env bash
# A/B of LPB_TUNE settings given in $VARIANTS (space separated), after the head tests
set -u
mkdir -p gpurun_out
timeout 500 python -m pytest tests -q -m gpu --timeout 120 -x -k "variants or head or tracker or mhcrnn or multiview or windows or hint or decode or predict" > gpurun_out/r2_tests_ab.log 2>&1
rc=$?; echo "head tests exit $rc"; tail -4 gpurun_out/r2_tests_ab.log
if [ $rc -ne 0 ]; then grep -n "Error\|error\|assert" gpurun_out/r2_tests_ab.log | head -20; exit 1; fi
for v in $VARIANTS; do
  LPB_TUNE="$v" timeout 400 python bench.py --steps 20 --warmup 5 --no-flat --no-cpu-baseline > "gpurun_out/r2_ab_${v//[=,]/_}.json" 2>/dev/null
  python - "$v" <<'PY'
import json,sys,re
v=sys.argv[1]
try:
    d=json.load(open("gpurun_out/r2_ab_%s.json"%re.sub("[=,]","_",v)))
    print(v, " ms/step", round(d["ms_per_step"],4), "fwd-only", round(d["forward_only"]["ms_per_step"],4), {k: s["ms"] for k,s in d["stages"].items() if k.startswith("head")})
except Exception as e: print(v, " failed", e)
PY
done
