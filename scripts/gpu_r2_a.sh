#!/usr/bin/env bash
# round 2, call A: new-shape parity tests first (bounded), then the whole GPU suite, smoke, a short bench
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu --timeout 240 -k "real_config or upsample2x or temporal_heatmap_loss_backward or fp32_native or one_deconv" > gpurun_out/r2a_new_tests.log 2>&1
echo "new tests exit $?" >> gpurun_out/r2a_new_tests.log
tail -25 gpurun_out/r2a_new_tests.log
timeout 1500 python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/r2a_all_tests.log 2>&1
echo "all tests exit $?" >> gpurun_out/r2a_all_tests.log
tail -15 gpurun_out/r2a_all_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/r2a_smoke.log
tail -3 gpurun_out/r2a_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench exit $?"
tail -c 3000 gpurun_out/r2a_bench.json
tail -5 gpurun_out/r2a_bench.err
