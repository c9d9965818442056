#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
# A: env var, no smoke before
AGR_STAGE_DETAIL=1 timeout -s ABRT -k 10 150 python -X faulthandler bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/h_a.json 2> gpurun_out/h_a.err; echo "rc=$?" >> gpurun_out/h_a.err
# B: smoke, then plain bench
( timeout -k 10 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 ) > gpurun_out/h_smoke.log
timeout -s ABRT -k 10 150 python -X faulthandler bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/h_b.json 2> gpurun_out/h_b.err; echo "rc=$?" >> gpurun_out/h_b.err
# C: default flags exactly as the driver runs it (cpu baseline included)
SECONDS=0
timeout -s ABRT -k 10 300 python -X faulthandler bench.py > gpurun_out/h_c.json 2> gpurun_out/h_c.err; echo "rc=$? wall=${SECONDS}s" >> gpurun_out/h_c.err
for f in a b c; do echo "== $f"; tail -c 1500 gpurun_out/h_$f.err | grep -v Warning | tail -25; cut -c1-120 gpurun_out/h_$f.json; done
