#!/bin/bash
# conv kernel development run: parity of generations 2 and 3 (tools/try_conv.py), device-time table, stock-reference diagnostic
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for gen in 2 3; do
  for grp in same down up view; do
    timeout -k 10 240 python tools/try_conv.py $grp $gen > gpurun_out/try_g${gen}_${grp}.log 2>&1; echo "rc=$?" >> gpurun_out/try_g${gen}_${grp}.log
  done
done
WGRAD_CTAS=74,148,296,592 timeout -k 10 400 python tools/bench_conv.py 1,2,3 > gpurun_out/bench_conv.log 2>&1; echo "rc=$?" >> gpurun_out/bench_conv.log
timeout -k 10 240 python tools/try_conv.py big 3 > gpurun_out/try_g3_big.log 2>&1; echo "rc=$?" >> gpurun_out/try_g3_big.log
timeout -k 10 600 python tools/diag_stock.py > gpurun_out/diag_stock.log 2>&1; echo "rc=$?" >> gpurun_out/diag_stock.log
tail -n 4 gpurun_out/try_g*.log; tail -5 gpurun_out/bench_conv.log; tail -3 gpurun_out/diag_stock.log
