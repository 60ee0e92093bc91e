#!/usr/bin/env bash
# memcheck of the kernels changed in the second session of round 2 (small shapes; bounded)
set -u
mkdir -p gpurun_out
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 77 --print-limit 20 \
  python -m pytest tests/test_gpu_parity.py -q -x --timeout 400 -k "fused_backward_vs_oracle or hinted_equals_plain or emits_decode_hints" > gpurun_out/r2_sanitize.log 2>&1
echo "sanitizer exit $?"
grep -c "Invalid\|Error:" gpurun_out/r2_sanitize.log; grep -m 12 "Invalid\|ERROR SUMMARY\|passed\|failed\|at .* in " gpurun_out/r2_sanitize.log | cut -c1-200
tail -5 gpurun_out/r2_sanitize.log | cut -c1-200
