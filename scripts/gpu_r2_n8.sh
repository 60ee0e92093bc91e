#!/usr/bin/env bash
# N-GPU bench with the driver's launch line (N from $1)
set -u
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 10 --warmup 3 --no-flat > gpurun_out/r2_n$N.json 2> gpurun_out/r2_n$N.err
echo "n$N exit $?"; grep -o '"value": [0-9.]*, "unit": "frames/s", "n_gpus": [0-9]*[^}]*"ms_per_step": [0-9.]*' gpurun_out/r2_n$N.json; grep -o '"ddp": "[^"]*"' gpurun_out/r2_n$N.json | cut -c1-200; tail -2 gpurun_out/r2_n$N.err
