#!/usr/bin/env bash
# decode warp kernel: occupancy cap x reverse order
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu -k "variants or decode" --timeout 300 2>&1 | tail -3
for v in "8=0" "8=6" "8=4" "8=3" "8=2" "9=1" "8=4,9=1" "8=3,9=1"; do
  LPB_TUNE="$v" timeout 400 python bench.py --steps 20 --warmup 5 --no-flat --no-cpu-baseline > "gpurun_out/r2_dec_${v//[=,]/_}.json" 2>/dev/null
  python - "$v" <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.load(open("gpurun_out/r2_dec_%s.json"%v.replace("=","_").replace(",","_")))
    print(v, " ms/step", round(d["ms_per_step"],4), "fwd-only", round(d["forward_only"]["ms_per_step"],4), "decode_fwd", d["stages"]["decode_fwd"]["ms"], "head_fwd", d["stages"]["head_fwd"]["ms"])
except Exception as e: print(v, " failed", e)
PY
done
