#!/bin/bash
# N-GPU checks: (1) one eager step at N=1 and N from the same seeded state: loss / gradient / parameter delta must agree
# (SURVEY.md §8e); (2) the bench line at N (view shard [+ owner-computes] + one all-reduce) and the clean NCCL teardown.
#   bash tools/gpu_multi.sh N [AGR_NET_PARALLEL value]
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${1:-2}
if [ -n "$2" ]; then export AGR_NET_PARALLEL=$2; fi
TAG="n${N}_np${AGR_NET_PARALLEL:-auto}"
timeout -k 10 300 python bench.py --check > gpurun_out/check_n1.json 2> gpurun_out/check_n1.err; echo "rc=$?" >> gpurun_out/check_n1.err
timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --check --check-against gpurun_out/check_n1.pt > gpurun_out/check_$TAG.json 2> gpurun_out/check_$TAG.err; echo "rc=$?" >> gpurun_out/check_$TAG.err
SECONDS=0
timeout -k 10 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "rc=$? wall=${SECONDS}s" >> gpurun_out/bench_$TAG.err
tail -n 1 gpurun_out/check_$TAG.json; tail -2 gpurun_out/check_$TAG.err | cut -c1-200; tail -n 1 gpurun_out/bench_$TAG.json | cut -c1-400; tail -2 gpurun_out/bench_$TAG.err | cut -c1-200
