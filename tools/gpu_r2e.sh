#!/bin/bash
# round-2 GPU call E (2 GPUs): distributed Cholesky over NCCL + C3 bench at N=2
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tools/check_dist_potrf.py 8192 50048 > $O/r2e_potrf.log 2>&1
echo "rc=$?" >> $O/r2e_potrf.log
timeout 900 $TR --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 1 --no-cpu-baseline --no-extras > $O/r2e_bench.log 2>&1
echo "rc=$?" >> $O/r2e_bench.log
tail -4 $O/r2e_potrf.log; tail -2 $O/r2e_bench.log | cut -c1-2500
