#!/bin/bash
# Round-end evidence run on one B200: full GPU test suite, smoke, headline bench line, launch list, the other BASELINE configs,
# the fp32 product arm, bf16-vs-fp32 full-size check, fresh ncu captures of the dominant conv shape.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out/ncu
( timeout -k 10 900 python -m pytest tests -m gpu -q 2>&1 | tail -25; echo "pytest_rc=${PIPESTATUS[0]}" ) > gpurun_out/f_pytest.log
( timeout -k 10 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4; echo "smoke_rc=${PIPESTATUS[0]}" ) > gpurun_out/f_smoke.log
AGR_STAGE_DETAIL=1 timeout -k 10 480 python bench.py --steps 10 --warmup 3 > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench_rc=$?" >> gpurun_out/f_bench.err
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/f_step_launches.csv python tools/prof_step.py 1 1 > gpurun_out/f_prof_step.log 2>&1
for c in 1 2 3; do timeout -k 10 300 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/f_bench_config$c.json 2> gpurun_out/f_bench_config$c.err; done
timeout -k 10 480 python bench.py --config 5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/f_bench_config5.json 2> gpurun_out/f_bench_config5.err
timeout -k 10 480 python bench.py --dtype fp32 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/f_bench_fp32.json 2> gpurun_out/f_bench_fp32.err
timeout -k 10 480 python bench.py --check-bf16 > gpurun_out/f_check_bf16.json 2> gpurun_out/f_check_bf16.err
N="ncu --set full --clock-control none --import-source on --profile-from-start off -f"
timeout -k 10 300 $N -k regex:conv_tc_kernel -c 1 -o gpurun_out/ncu/r02b_conv64 python tools/ncu_targets.py conv64 > gpurun_out/ncu/b_conv64.log 2>&1
timeout -k 10 300 $N -k regex:conv_tc3_kernel -c 1 -o gpurun_out/ncu/r02b_conv64_persistent python tools/ncu_targets.py conv64p > gpurun_out/ncu/b_conv64p.log 2>&1
timeout -k 10 300 $N -k regex:conv_wgrad_tc_kernel -c 1 -o gpurun_out/ncu/r02b_wgrad64 python tools/ncu_targets.py wgrad64 > gpurun_out/ncu/b_wgrad64.log 2>&1
tail -3 gpurun_out/f_pytest.log; tail -2 gpurun_out/f_smoke.log; cut -c1-260 gpurun_out/f_bench.json; for c in 1 2 3 5; do cut -c1-200 gpurun_out/f_bench_config$c.json; done; cut -c1-200 gpurun_out/f_bench_fp32.json; cat gpurun_out/f_check_bf16.json | cut -c1-600
