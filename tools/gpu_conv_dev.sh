#!/bin/bash
# conv kernel development run: parity (tools/try_conv.py) per generation, device-time table
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for gen in 1 3; do
  for grp in same down up view; do
    timeout -k 10 240 python tools/try_conv.py $grp $gen > gpurun_out/try_g${gen}_${grp}.log 2>&1; echo "rc=$?" >> gpurun_out/try_g${gen}_${grp}.log
  done
done
AGR_CONV_PERSISTENT=1 timeout -k 10 300 python tools/try_conv.py big 0 > gpurun_out/try_g0p_big.log 2>&1; echo "rc=$?" >> gpurun_out/try_g0p_big.log
timeout -k 10 400 python tools/bench_conv.py 1,2,3 > gpurun_out/bench_conv3.log 2>&1; echo "rc=$?" >> gpurun_out/bench_conv3.log
AGR_CONV_PERSISTENT=1 timeout -k 10 400 python tools/bench_conv.py 0 > gpurun_out/bench_conv3p.log 2>&1; echo "rc=$?" >> gpurun_out/bench_conv3p.log
grep -c FAIL gpurun_out/try_g*.log; tail -n 2 gpurun_out/try_g*.log | grep -v "^$"; tail -3 gpurun_out/bench_conv3.log; tail -2 gpurun_out/bench_conv3p.log
