#!/bin/bash
# One-shot GPU validation: GPU parity tests, smoke(), a short bench, and an ncu launch list of one eager step.
# Outputs land in gpurun_out/.  (gpurun --timeout 900 -- 'bash tools/validate_gpu.sh')
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25; echo "pytest_rc=${PIPESTATUS[0]}" ) > gpurun_out/v_pytest.log
( timeout 150 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5; echo "smoke_rc=${PIPESTATUS[0]}" ) > gpurun_out/v_smoke.log
AGR_STAGE_DETAIL=1 timeout 480 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/v_bench.json 2> gpurun_out/v_bench.err; echo "bench_rc=$?" >> gpurun_out/v_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/step_launches.csv python tools/prof_step.py 1 1 > gpurun_out/prof_step.log 2>&1
tail -3 gpurun_out/v_pytest.log; tail -2 gpurun_out/v_smoke.log; cut -c1-300 gpurun_out/v_bench.json; tail -2 gpurun_out/v_bench.err; wc -l gpurun_out/step_launches.csv
timeout -k 10 300 python tools/try_conv.py big 0 > gpurun_out/try_g0_big.log 2>&1; tail -2 gpurun_out/try_g0_big.log
timeout -k 10 300 python tools/bench_conv.py 1,0 > gpurun_out/bench_conv4.log 2>&1; tail -2 gpurun_out/bench_conv4.log
