#!/usr/bin/env bash
# dgrad zero-chunk skip: parity, bench, launch list
set -u
mkdir -p gpurun_out
timeout 500 python -m pytest tests -q -m gpu --timeout 120 -x -k "variants or head or tracker or adam or mhcrnn or multiview" > gpurun_out/r2_tests_c2.log 2>&1
rc=$?; echo "head tests exit $rc"; tail -6 gpurun_out/r2_tests_c2.log
if [ $rc -ne 0 ]; then grep -n "Error\|error\|assert" gpurun_out/r2_tests_c2.log | head -20; fi
timeout 400 python bench.py --steps 20 --warmup 5 --no-flat --no-cpu-baseline > gpurun_out/r2_c2.json 2>/dev/null
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r2_c2.json"))
    print(" ms/step", round(d["ms_per_step"],4), "fwd", round(d["forward_only"]["ms_per_step"],4), {k: s["ms"] for k,s in d["stages"].items()})
except Exception as e: print(" failed", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_c2_launches.csv \
  python bench.py --profile-step --steps 2 --warmup 1 --no-graph --serial-chains > gpurun_out/r2_c2_ncu.log 2>&1
echo "ncu exit $?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/r2_c2_launches.csv")) if len(r)>5]
hdr=rows[0]; ix={h:i for i,h in enumerate(hdr)}
body=[r for r in rows[1:] if r[ix["Metric Name"]]=="gpu__time_duration.sum"]
n=len(body); half=body[n//2:]  # second step
tot=0
for r in half:
    v=float(r[ix["Metric Value"]].replace(",","")); u=r[ix["Metric Unit"]]
    us = v/1000 if u in ("ns","nsecond") else v
    tot+=us
    print("%-60s %8.1f"%(r[ix["Kernel Name"]][:60], us))
print("total", tot)
PY
