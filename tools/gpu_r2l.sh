#!/bin/bash
# round-2 GPU call L (2 GPUs): full default bench at N=2 with the final code (pinned staging, LL solves)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29532 bench.py --gpus $N --no-cpu-baseline > $O/r2l_bench_n$N.log 2>&1
echo "rc=$?" >> $O/r2l_bench_n$N.log
timeout 600 python -m pytest tests/test_gpu_fit.py -q -m gpu -k "sharded" > $O/r2l_tests.log 2>&1
tail -2 $O/r2l_bench_n$N.log | cut -c1-2500; tail -2 $O/r2l_tests.log
