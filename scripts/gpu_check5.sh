#!/usr/bin/env bash
# decode tests on both forward paths, then the usual check
set -uo pipefail
mkdir -p gpurun_out
echo "== decode tests, CTA-only path"; LPB_DECODE_CTA_ONLY=1 timeout 300 python -m pytest tests -m gpu -x -q -k "decode or tracker or keypoints" 2>&1 | tail -3 | tee gpurun_out/pytest_cta.log
echo "== decode tests, warp path"; timeout 300 python -m pytest tests -m gpu -x -q -k "decode or tracker or keypoints" 2>&1 | tail -15 | tee gpurun_out/pytest_warp.log
if grep -q "failed\|rror" gpurun_out/pytest_warp.log; then echo "ABORT"; exit 1; fi
bash scripts/gpu_check4.sh
