#!/usr/bin/env bash
# quick: tests + bench + launch list of one step
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 300 -x > gpurun_out/r2_tests.log 2>&1
echo "tests exit $?"; tail -4 gpurun_out/r2_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-flat > gpurun_out/r2_bench_b7.json 2> gpurun_out/r2_bench_b7.err
echo "bench exit $?"; python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r2_bench_b7.json"))
    print(" ms/step", round(d["ms_per_step"],4), "fwd-only", round(d["forward_only"]["ms_per_step"],4))
    print(" stages", {k: v["ms"] for k,v in d["stages"].items()})
except Exception as e: print(" failed", e)
PY
tail -3 gpurun_out/r2_bench_b7.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_b7.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-flat --no-graph --serial-chains > gpurun_out/r02_ncu_b7.log 2>&1
echo "ncu exit $?"
