#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
: > $O/sweep_trsm.log
for c in 256 384 512 1024; do
  echo "-- SAB_TRSM_CHUNK=$c" >> $O/sweep_trsm.log
  SAB_TRSM_CHUNK=$c BENCH_M=8192,16384,50048,100096 timeout 600 python tools/gpu_check.py bench_linalg 2>&1 | grep "M=" | sed 's/potrf.*pack/pack/' >> $O/sweep_trsm.log
done
cat $O/sweep_trsm.log
