#!/usr/bin/env bash
# fwd bring-up -> bwd bring-up -> pytest (abort early on any failure)
set -uo pipefail
mkdir -p gpurun_out
echo "== tcgen05 head bring-up"; HB_B=5 timeout 120 python scripts/test_head_bf16.py 2>&1 | tail -6 | tee gpurun_out/head_bf16.log
if ! grep -q "RESULT PASS" gpurun_out/head_bf16.log || grep -q "RESULT FAIL" gpurun_out/head_bf16.log; then echo "ABORT: tcgen05 head failed or hung"; exit 1; fi
echo "== head bwd bring-up"; timeout 200 python scripts/test_head_bwd_bf16.py 2>&1 | tail -8 | tee gpurun_out/bwd_test.log
if ! grep -q "ALL PASS" gpurun_out/bwd_test.log; then echo "ABORT: head backward failed or hung"; exit 1; fi
echo "== pytest gpu"; timeout 400 python -m pytest tests -m gpu -x -q --durations=6 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
if ! grep -q " passed" gpurun_out/pytest_gpu.log || grep -q "failed" gpurun_out/pytest_gpu.log; then echo "ABORT: gpu tests failed"; exit 1; fi
echo "== bench train"; timeout 400 python bench.py --steps 10 --warmup 3 ${BENCH_ARGS:-} 2>&1 | tail -3 | tee gpurun_out/bench.log
if [[ "${QUICK:-0}" == "1" ]]; then exit 0; fi
echo "== ncu launch list (train step)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_train.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-flat --no-graph > gpurun_out/ncu_list_train.log 2>&1
tail -1 gpurun_out/ncu_list_train.log | cut -c1-200
