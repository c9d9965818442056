#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout -k 10 200 python tools/dbg_lpips.py > gpurun_out/dbg_lpips.log 2>&1; echo "rc=$?" >> gpurun_out/dbg_lpips.log
timeout -s ABRT -k 10 200 python -X faulthandler bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/dbg_bench_a.json 2> gpurun_out/dbg_bench_a.err; echo "rc=$?" >> gpurun_out/dbg_bench_a.err
AGR_CONV_TC=1 timeout -s ABRT -k 10 200 python -X faulthandler bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/dbg_bench_b.json 2> gpurun_out/dbg_bench_b.err; echo "rc=$?" >> gpurun_out/dbg_bench_b.err
timeout -s ABRT -k 10 240 python -X faulthandler bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/dbg_bench_c.json 2> gpurun_out/dbg_bench_c.err; echo "rc=$?" >> gpurun_out/dbg_bench_c.err
( timeout -k 10 400 python -m pytest tests/test_reference_stock.py -q -x -s 2>&1 | tail -40 ) > gpurun_out/dbg_stock.log
tail -5 gpurun_out/dbg_lpips.log; for f in a b c; do tail -c 600 gpurun_out/dbg_bench_$f.err; cut -c1-150 gpurun_out/dbg_bench_$f.json; done
