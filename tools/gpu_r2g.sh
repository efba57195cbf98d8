#!/bin/bash
# round-2 GPU call G (8 GPUs): distributed Cholesky over NCCL + the default bench at N=8 (C3 headline, stored blocks, C4 leg)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29521 tools/check_dist_potrf.py 50048 > $O/r2g_potrf_n$N.log 2>&1
echo "rc=$?" >> $O/r2g_potrf_n$N.log
timeout 800 $TR --master-port 29522 bench.py --gpus $N --no-cpu-baseline > $O/r2g_bench_n$N.log 2>&1
echo "rc=$?" >> $O/r2g_bench_n$N.log
grep "M=" $O/r2g_potrf_n$N.log; tail -2 $O/r2g_bench_n$N.log | cut -c1-3000
