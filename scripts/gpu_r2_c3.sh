#!/usr/bin/env bash
# in-situ kernel durations (torch.profiler / CUPTI) of the eager step, window patch pass vs fused look-ups
set -u
mkdir -p gpurun_out
for v in "12=1" "12=0"; do
  LPB_TUNE="$v" timeout 300 python bench.py --profile-step --kineto --steps 5 --warmup 3 --no-graph --serial-chains > "gpurun_out/r2_kineto_${v//[=,]/_}.json" 2> gpurun_out/r2_kineto.err
  echo "exit $?"; tail -2 gpurun_out/r2_kineto.err
  python - "$v" <<'PY'
import json,sys,re
v=sys.argv[1]
d=json.load(open("gpurun_out/r2_kineto_%s.json"%re.sub("[=,]","_",v)))
k=d["kineto"]
if isinstance(k,dict) and "last_step" in k:
    tot=0
    for n,us in k["last_step"]:
        tot+=us
        if us>=5: print("%-62s %8.1f"%(n,us))
    print(v,"total",round(tot,1),"launches",len(k["last_step"]))
else:
    print(k)
PY
done
