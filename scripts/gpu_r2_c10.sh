#!/usr/bin/env bash
# full capture of k1a (training form, 512-frame launch) with source counters
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k1a_shuffle' -s 3 -c 1 \
  -o gpurun_out/r2_k1a python bench.py --profile-step --steps 2 --warmup 1 --no-graph --serial-chains > gpurun_out/r2_k1a_ncu.log 2>&1
echo "ncu exit $?"
ncu -i gpurun_out/r2_k1a.ncu-rep --page details 2>/dev/null | grep -E "Duration|Throughput|Issued Warp|Eligible|Active Warps|Warp Cycles Per Issued|stalled|Est. Speedup|Registers Per|Achieved Occupancy|Bank|L2 Hit|DRAM Throughput|Pipe" | head -40
ncu -i gpurun_out/r2_k1a.ncu-rep --page source --csv > gpurun_out/r2_k1a_src.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open("gpurun_out/r2_k1a_src.csv")))
hdr=rows[0]
print(hdr[:12])
ix={h:i for i,h in enumerate(hdr)}
samp=[h for h in hdr if "Sampl" in h][:3]
print(samp)
key=samp[0] if samp else None
body=rows[1:]
def f(x):
    try: return float(x.replace(",",""))
    except: return 0.0
tot=sum(f(r[ix[key]]) for r in body)
top=sorted(body,key=lambda r:-f(r[ix[key]]))[:40]
for r in top:
    print(round(100*f(r[ix[key]])/tot,1), r[ix.get("Address",0)][-6:], r[ix["Source"]][:110] if "Source" in ix else r[1][:110])
PY
