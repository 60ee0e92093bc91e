#!/usr/bin/env bash
# tcgen05 head bring-up (guarded by timeouts), then ncu launch list + full captures.
set -uo pipefail
mkdir -p gpurun_out
for swap in 0 1; do
  echo "== head bf16 swap=$swap"; LPB_DESC_SWAP=$swap timeout 120 python scripts/test_head_bf16.py 2>&1 | tail -8 | tee gpurun_out/head_bf16_swap$swap.log
  echo "rc=$?"
done
nvidia-smi --query-gpu=name,memory.used --format=csv
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 ${BENCH_ARGS:-} 2>&1 | tail -3 | tee gpurun_out/bench.log
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-flat > gpurun_out/ncu_list.log 2>&1
tail -3 gpurun_out/ncu_list.log
echo "== ncu full (decode, head)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"decode_fwd|convt3x3s2|k1a|k1b" -c 6 -o gpurun_out/prof_r01 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-flat > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
