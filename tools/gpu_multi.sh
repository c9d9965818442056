#!/bin/bash
# 2-GPU checks: (1) one eager step at N=1 and N=2 from the same seeded state: loss / gradient / parameter delta must agree
# (SURVEY.md §8e); (2) the bench line at N=2 (view shard + one all-reduce) and the clean NCCL teardown.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${1:-2}
timeout -k 10 300 python bench.py --check > gpurun_out/check_n1.json 2> gpurun_out/check_n1.err; echo "rc=$?" >> gpurun_out/check_n1.err
timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --check --check-against gpurun_out/check_n1.pt > gpurun_out/check_n$N.json 2> gpurun_out/check_n$N.err; echo "rc=$?" >> gpurun_out/check_n$N.err
SECONDS=0
timeout -k 10 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$? wall=${SECONDS}s" >> gpurun_out/bench_n$N.err
cat gpurun_out/check_n1.json; cat gpurun_out/check_n$N.json; tail -2 gpurun_out/check_n$N.err; cut -c1-400 gpurun_out/bench_n$N.json; tail -3 gpurun_out/bench_n$N.err
