#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for grp in same up; do timeout -k 10 200 python tools/try_conv.py $grp 0 > gpurun_out/s2_try_$grp.log 2>&1; tail -1 gpurun_out/s2_try_$grp.log; done
AGR_CONV_PERSISTENT=1 timeout -k 10 200 python tools/try_conv.py big 0 > gpurun_out/s2_try_big_p.log 2>&1; tail -1 gpurun_out/s2_try_big_p.log
( timeout -k 10 400 python -m pytest tests/test_styleunet.py tests/test_avatar.py -m gpu -q 2>&1 | tail -3 ) > gpurun_out/s2_pytest.log; tail -1 gpurun_out/s2_pytest.log
ROUNDS=6 MODES="default pers1" bash tools/gpu_stress.sh
timeout -k 10 300 python tools/bench_conv.py 1,0 > gpurun_out/bench_conv6.log 2>&1; tail -1 gpurun_out/bench_conv6.log
