#!/usr/bin/env bash
# 2-GPU bench (driver's launch line) with the two-stream chains + the 94 MB bucket
set -u
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 --no-flat > gpurun_out/r2_n2.json 2> gpurun_out/r2_n2.err
echo "n2 exit $?"; tail -c 1200 gpurun_out/r2_n2.json | head -c 400; echo; grep -o '"value": [0-9.]*, "unit": "frames/s", "n_gpus": 2[^}]*"ms_per_step": [0-9.]*' gpurun_out/r2_n2.json; tail -3 gpurun_out/r2_n2.err
