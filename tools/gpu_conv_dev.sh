#!/bin/bash
# conv kernel development run: parity (tools/try_conv.py) with the shape-based kernel choice, device-time table, weight-gradient
# split sweep, then the full GPU test suite, a bench line and the stock-reference diagnostic
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for grp in same down up view big; do
  timeout -k 10 300 python tools/try_conv.py $grp 0 > gpurun_out/try_g0_${grp}.log 2>&1; echo "rc=$?" >> gpurun_out/try_g0_${grp}.log
done
WGRAD_SPLIT=592:16,592:32,592:8,296:16,1184:16 timeout -k 10 400 python tools/bench_conv.py 1,0 > gpurun_out/bench_conv2.log 2>&1; echo "rc=$?" >> gpurun_out/bench_conv2.log
( timeout -k 10 700 python -m pytest tests -m gpu -q -x 2>&1 | tail -25; echo "pytest_rc=${PIPESTATUS[0]}" ) > gpurun_out/v_pytest.log
AGR_STAGE_DETAIL=1 timeout -k 10 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/v_bench.json 2> gpurun_out/v_bench.err; echo "bench_rc=$?" >> gpurun_out/v_bench.err
for img in 512 1024; do DIAG_IMG=$img timeout -k 10 400 python tools/diag_stock.py > gpurun_out/diag_stock_$img.log 2>&1; echo "rc=$?" >> gpurun_out/diag_stock_$img.log; done
grep -c FAIL gpurun_out/try_g0_*.log; tail -n 2 gpurun_out/try_g0_*.log; tail -3 gpurun_out/bench_conv2.log; tail -3 gpurun_out/v_pytest.log; cut -c1-200 gpurun_out/v_bench.json
