#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
( timeout -k 10 500 python -m pytest tests/test_styleunet.py tests/test_avatar.py tests/test_lpips.py -m gpu -q 2>&1 | tail -8 ) > gpurun_out/ab_pytest.log
SECONDS=0
timeout -k 10 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ab_bench_par.json 2> gpurun_out/ab_bench_par.err; echo "rc=$? wall=${SECONDS}s" >> gpurun_out/ab_bench_par.err
AGR_SERIAL_DECODERS=1 timeout -k 10 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ab_bench_ser.json 2> gpurun_out/ab_bench_ser.err; echo "rc=$?" >> gpurun_out/ab_bench_ser.err
AGR_CONV_PERSISTENT=1 timeout -k 10 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ab_bench_pers.json 2> gpurun_out/ab_bench_pers.err; echo "rc=$?" >> gpurun_out/ab_bench_pers.err
AGR_CONV_PERSISTENT=1 timeout -k 10 200 python tools/bench_conv.py 0 > gpurun_out/bench_conv5p.log 2>&1
tail -4 gpurun_out/ab_pytest.log; for f in par ser pers; do tail -1 gpurun_out/ab_bench_$f.err; python -c "
import json;d=json.load(open('gpurun_out/ab_bench_$f.json')); print('$f', round(d['value'],1), round(d['ms_per_step'],2), round(d['e2e']['value'],1))"; done; head -5 gpurun_out/bench_conv5p.log
