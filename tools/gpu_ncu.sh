#!/bin/bash
# ncu --set full captures of the kernels VERDICT / DESIGN discuss (one launch each); reports land in gpurun_out/ncu/
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out/ncu
N="ncu --set full --clock-control none --import-source on --profile-from-start off -f"
timeout -k 10 300 $N -k regex:conv_tc_kernel -c 1 -o gpurun_out/ncu/r02_conv64 python tools/ncu_targets.py conv64 > gpurun_out/ncu/conv64.log 2>&1
timeout -k 10 300 $N -k regex:conv_tc3_kernel -c 1 -o gpurun_out/ncu/r02_conv64_persistent python tools/ncu_targets.py conv64p > gpurun_out/ncu/conv64p.log 2>&1
timeout -k 10 300 $N -k regex:conv_tc_kernel -c 1 -o gpurun_out/ncu/r02_conv128 python tools/ncu_targets.py conv128 > gpurun_out/ncu/conv128.log 2>&1
timeout -k 10 300 $N -k regex:conv_wgrad_tc_kernel -c 1 -o gpurun_out/ncu/r02_wgrad64 python tools/ncu_targets.py wgrad64 > gpurun_out/ncu/wgrad64.log 2>&1
timeout -k 10 300 $N -k regex:lbs_ -c 2 -o gpurun_out/ncu/r02_lbs python tools/ncu_targets.py lbs > gpurun_out/ncu/lbs.log 2>&1
timeout -k 10 400 $N -k regex:blend_ -c 2 -o gpurun_out/ncu/r02_blend python tools/ncu_targets.py raster > gpurun_out/ncu/raster.log 2>&1
ls -la gpurun_out/ncu; tail -2 gpurun_out/ncu/*.log
