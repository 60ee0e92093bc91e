#!/usr/bin/env bash
# last check of the round on the final commit: full GPU suite + smoke
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/r2_tests.log 2>&1
echo "tests exit $?"; tail -3 gpurun_out/r2_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1
echo "smoke exit $?"; tail -2 gpurun_out/r2_smoke.log
