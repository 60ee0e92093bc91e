#!/usr/bin/env bash
# GPU parity + bench + ncu launch list + full captures of the top kernels.
# Risky kernels first, under short timeouts: a hang costs ~2 minutes, not the whole budget.
set -uo pipefail
mkdir -p gpurun_out
echo "== first-call timings"; timeout 300 python scripts/time_first_calls.py 2>&1 | tail -20 | tee gpurun_out/first_calls.log
echo "== tcgen05 head bring-up"; HB_B=5 timeout 120 python scripts/test_head_bf16.py 2>&1 | tail -6 | tee gpurun_out/head_bf16.log
if ! grep -q "RESULT PASS" gpurun_out/head_bf16.log || grep -q "RESULT FAIL" gpurun_out/head_bf16.log; then echo "ABORT: tcgen05 head failed or hung"; exit 1; fi
echo "== pytest gpu"; timeout 400 python -m pytest tests -m gpu -x -q --durations=6 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
if ! grep -q " passed" gpurun_out/pytest_gpu.log || grep -q "failed" gpurun_out/pytest_gpu.log; then echo "ABORT: gpu tests failed"; exit 1; fi
echo "== bench bf16"; timeout 400 python bench.py --steps 10 --warmup 3 ${BENCH_ARGS:-} 2>&1 | tail -3 | tee gpurun_out/bench.log
if [[ "${QUICK:-0}" == "1" ]]; then exit 0; fi
echo "== bench f32"; timeout 400 python bench.py --steps 5 --warmup 3 --dtype f32 --no-cpu-baseline 2>&1 | tail -3 | tee gpurun_out/bench_f32.log
echo "== ncu launch list"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-flat > gpurun_out/ncu_list.log 2>&1
tail -2 gpurun_out/ncu_list.log | cut -c1-300
echo "== ncu full"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"decode_fwd|k1a|k1b" -c 6 -o gpurun_out/prof_r01b -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-flat > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-300
ls -la gpurun_out
