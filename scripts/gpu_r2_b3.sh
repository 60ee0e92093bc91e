#!/usr/bin/env bash
# experiments: two-stream chains A/B, fresh-init regime launch list
set -u
mkdir -p gpurun_out
for mode in "" "--serial-chains"; do
  tag=${mode:+serial}; tag=${tag:-forked}
  timeout 600 python bench.py --steps 20 --warmup 5 --no-flat --no-cpu-baseline $mode > gpurun_out/r2_chains_$tag.json 2> gpurun_out/r2_chains_$tag.err
  echo "$tag exit $?"; python - "$tag" <<'PY'
import json,sys
try:
    d=json.load(open("gpurun_out/r2_chains_%s.json"%sys.argv[1]))
    print(" ms/step", round(d["ms_per_step"],4), "fwd-only", round(d["forward_only"]["ms_per_step"],4), "e2e", round(d["e2e"]["value"]))
except Exception as e: print(" failed", e)
PY
  tail -3 gpurun_out/r2_chains_$tag.err
done
timeout 600 python bench.py --steps 5 --warmup 3 --regime fresh --no-flat --no-cpu-baseline > gpurun_out/r2_fresh.json 2> gpurun_out/r2_fresh.err
echo "fresh exit $?"; tail -c 1500 gpurun_out/r2_fresh.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_fresh.csv \
  python bench.py --steps 1 --warmup 1 --regime fresh --no-cpu-baseline --no-flat --no-graph > gpurun_out/r02_ncu_fresh.log 2>&1
echo "ncu fresh exit $?"
