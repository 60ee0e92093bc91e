#!/usr/bin/env bash
# full ncu capture of the dense decode kernels in the fresh-init regime
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'decode_fwd_kernel|decode_bwd_kernel' -s 4 -c 4 \
  -o gpurun_out/r02_dense python bench.py --steps 1 --warmup 1 --regime fresh --no-cpu-baseline --no-flat --no-graph > gpurun_out/r02_ncu_dense.log 2>&1
echo "ncu exit $?"; tail -5 gpurun_out/r02_ncu_dense.log
ncu -i gpurun_out/r02_dense.ncu-rep --page raw --csv > gpurun_out/r02_dense_raw.csv 2>/dev/null
ncu -i gpurun_out/r02_dense.ncu-rep --page source --csv --kernel-name regex:decode_fwd_kernel --launch-skip 0 --launch-count 1 > gpurun_out/r02_dense_src_fwd.csv 2>/dev/null
ncu -i gpurun_out/r02_dense.ncu-rep --page source --csv --kernel-name regex:decode_bwd_kernel --launch-skip 0 --launch-count 1 > gpurun_out/r02_dense_src_bwd.csv 2>/dev/null
ls -la gpurun_out | grep dense
