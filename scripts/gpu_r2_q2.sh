#!/usr/bin/env bash
# quick: head-related tests (short timeouts) + one bench line
set -u
mkdir -p gpurun_out
timeout 500 python -m pytest tests -q -m gpu --timeout 120 -x -k "variants or head or tracker or adam or mhcrnn or multiview or windows or predict" > gpurun_out/r2_tests_q2.log 2>&1
rc=$?; echo "head tests exit $rc"; tail -6 gpurun_out/r2_tests_q2.log
if [ $rc -ne 0 ]; then grep -n "Error\|error\|assert" gpurun_out/r2_tests_q2.log | head -20; exit 1; fi
LPB_TUNE="${TUNE:-}" timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline ${BENCH_FLAGS:-} > gpurun_out/r2_q2.json 2>gpurun_out/r2_q2.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r2_q2.json"))
    print(" ms/step", round(d["ms_per_step"],4), "fwd-only", round(d["forward_only"]["ms_per_step"],4), {k: s["ms"] for k,s in d["stages"].items()})
    f=d.get("fresh_init_regime")
    if f: print("   fresh", round(f["value"]), f["stages"])
except Exception as e: print(" failed", e)
PY
tail -3 gpurun_out/r2_q2.err
