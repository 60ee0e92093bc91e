#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 500 python -m pytest tests -q -m gpu --timeout 120 -x -k "variants or head or tracker or adam or mhcrnn or multiview or windows" > gpurun_out/r2_tests_c6.log 2>&1
rc=$?; echo "head tests exit $rc"; tail -6 gpurun_out/r2_tests_c6.log
if [ $rc -ne 0 ]; then grep -n "Error\|error\|assert" gpurun_out/r2_tests_c6.log | head -20; exit 1; fi
for v in "12=1" "12=0"; do
  LPB_TUNE="$v" timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > "gpurun_out/r2_c6_${v//[=,]/_}.json" 2>/dev/null
  python - "$v" <<'PY'
import json,sys,re
v=sys.argv[1]
try:
    d=json.load(open("gpurun_out/r2_c6_%s.json"%re.sub("[=,]","_",v)))
    print(v, " ms/step", round(d["ms_per_step"],4), "bwd_lab", d["stages"]["head_bwd_labeled"]["ms"], "bwd_unl", d["stages"]["head_bwd_unlabeled"]["ms"])
    f=d.get("fresh_init_regime")
    if f: print("   fresh", round(f["value"]), f["stages"])
except Exception as e: print(v, " failed", e)
PY
done
LPB_TUNE="12=1" timeout 300 python bench.py --profile-step --kineto --steps 5 --warmup 3 --no-graph --serial-chains > "gpurun_out/r2_kineto_c6.json" 2> gpurun_out/r2_kineto.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_kineto_c6.json"))
k=d["kineto"]
tot=0
for n,us in k["last_step"]:
    tot+=us
    if us>=10 and ("g2_" in n or "plane_dot" in n or "prep" in n): print("%-62s %8.1f"%(n,us))
print("total",round(tot,1),"launches",len(k["last_step"]))
PY
