#!/bin/bash
# One-shot GPU validation: GPU parity tests, smoke(), a short bench.  Outputs land in gpurun_out/.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
( timeout 420 python -m pytest tests -m gpu -x -q 2>&1 | tail -25; echo "pytest_rc=${PIPESTATUS[0]}" ) > gpurun_out/v_pytest.log
( timeout 150 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5; echo "smoke_rc=${PIPESTATUS[0]}" ) > gpurun_out/v_smoke.log
timeout 240 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/v_bench.json 2> gpurun_out/v_bench.err; echo "bench_rc=$?" >> gpurun_out/v_bench.err
tail -3 gpurun_out/v_pytest.log; tail -2 gpurun_out/v_smoke.log; cut -c1-400 gpurun_out/v_bench.json; tail -3 gpurun_out/v_bench.err
