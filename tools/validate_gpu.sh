#!/bin/bash
# One-shot GPU validation: GPU parity tests, a short bench, and an ncu launch list of one eager step.
# Outputs land in gpurun_out/.  If the tests fail with split-K convolutions on, everything is repeated with them off.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run_tests() { ( timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -25; echo "pytest_rc=${PIPESTATUS[0]}" ) > "$1"; grep -q "pytest_rc=0" "$1"; }
export AGR_CONV_SPLITK=1
if ! run_tests gpurun_out/v_pytest.log; then
  export AGR_CONV_SPLITK=0
  run_tests gpurun_out/v_pytest_nosplit.log
fi
echo "AGR_CONV_SPLITK=$AGR_CONV_SPLITK" > gpurun_out/v_mode.txt
timeout 240 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/v_bench.json 2> gpurun_out/v_bench.err; echo "bench_rc=$?" >> gpurun_out/v_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/step_launches_c.csv python tools/prof_step.py 1 1 > gpurun_out/prof_step_c.log 2>&1
cat gpurun_out/v_mode.txt; tail -3 gpurun_out/v_pytest.log; [ -f gpurun_out/v_pytest_nosplit.log ] && tail -3 gpurun_out/v_pytest_nosplit.log
cut -c1-300 gpurun_out/v_bench.json; tail -2 gpurun_out/v_bench.err; wc -l gpurun_out/step_launches_c.csv
