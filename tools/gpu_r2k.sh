#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "cholesky or paired or large_fact" > $O/r2k_tests.log 2>&1
echo "rc=$?" >> $O/r2k_tests.log
timeout 300 python tools/trace_trsm.py 50048 20 0 > $O/r2k_trace.log 2>&1
timeout 300 python tools/trace_trsm.py 50048 20 1 >> $O/r2k_trace.log 2>&1
BENCH_M=16384,50048,100096 timeout 600 python tools/gpu_check.py bench_linalg > $O/r2k_linalg.log 2>&1
grep -E "passed|failed|rc=" $O/r2k_tests.log; cat $O/r2k_trace.log | head -9; grep "M=" $O/r2k_linalg.log
