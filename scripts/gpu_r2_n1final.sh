#!/usr/bin/env bash
# final N=1 bench line of the round (default flags as the driver runs it)
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > gpurun_out/r02_gpu_box.txt 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_train.json 2> gpurun_out/r02_bench_train.err
echo "bench exit $?"; python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_bench_train.json"))
print("ms/step",round(d["ms_per_step"],4),"value",round(d["value"]),"e2e",round(d["e2e"]["value"]),"fwd",round(d["forward_only"]["ms_per_step"],4),"frac",round(d["roofline"]["frac"],3))
print({k:s["ms"] for k,s in d["stages"].items()})
print("fresh",round(d["fresh_init_regime"]["value"]))
PY
