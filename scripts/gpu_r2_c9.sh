#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 500 python -m pytest tests -q -m gpu --timeout 120 -x -k "variants or head or hint or tracker or mhcrnn or multiview" > gpurun_out/r2_tests_c11.log 2>&1
rc=$?; echo "tests exit $rc"; tail -4 gpurun_out/r2_tests_c11.log
if [ $rc -ne 0 ]; then grep -n "Error\|error\|assert" gpurun_out/r2_tests_c11.log | head -20; exit 1; fi
for v in "15=2"; do
  LPB_TUNE="$v" timeout 400 python bench.py --steps 20 --warmup 5 --no-flat --no-cpu-baseline > "gpurun_out/r2_c11_${v//[=,]/_}.json" 2>/dev/null
  python - "$v" <<'PY'
import json,sys,re
v=sys.argv[1]
try:
    d=json.load(open("gpurun_out/r2_c11_%s.json"%re.sub("[=,]","_",v)))
    print(v, " ms/step", round(d["ms_per_step"],4), "fwd-only", round(d["forward_only"]["ms_per_step"],4), {k: s["ms"] for k,s in d["stages"].items()})
except Exception as e: print(v, " failed", e)
PY
done
timeout 300 python bench.py --profile-step --kineto --steps 5 --warmup 3 --no-graph --serial-chains > "gpurun_out/r2_kineto_c11.json" 2> gpurun_out/r2_kineto.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_kineto_c11.json"))
k=d["kineto"]
tot=0
for n,us in k["last_step"]:
    tot+=us
    if us>=20 and ("convt_rows" in n or "decode_fwd" in n or "k1a" in n): print("%-62s %8.1f"%(n,us))
print("total",round(tot,1),"launches",len(k["last_step"]))
PY
