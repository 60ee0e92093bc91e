#!/usr/bin/env bash
# quick: tests + bench (no ncu)
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 300 -x > gpurun_out/r2_tests.log 2>&1
echo "tests exit $?"; tail -4 gpurun_out/r2_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline ${BENCH_FLAGS:-} > gpurun_out/r2_bench_q.json 2> gpurun_out/r2_bench_q.err
echo "bench exit $?"; python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r2_bench_q.json"))
    print(" ms/step", round(d["ms_per_step"],4), "fwd-only", round(d["forward_only"]["ms_per_step"],4), "e2e", round(d["e2e"]["value"]))
    print(" stages", {k: v["ms"] for k,v in d["stages"].items()})
    if d.get("fresh_init_regime"): print(" fresh", round(d["fresh_init_regime"]["value"]), d["fresh_init_regime"]["stages"])
except Exception as e: print(" failed", e)
PY
tail -3 gpurun_out/r2_bench_q.err
