#!/bin/bash
# Short A/B: StyleUNet GPU tests + bench with split-K convolutions on.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export AGR_CONV_SPLITK=1
( timeout 100 python -m pytest tests/test_styleunet.py -m gpu -q 2>&1 | tail -6; echo "pytest_rc=${PIPESTATUS[0]}" ) > gpurun_out/w_pytest.log
timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/w_bench_splitk.json 2> gpurun_out/w_bench.err; echo "bench_rc=$?" >> gpurun_out/w_bench.err
tail -3 gpurun_out/w_pytest.log; cut -c1-300 gpurun_out/w_bench_splitk.json; tail -2 gpurun_out/w_bench.err
