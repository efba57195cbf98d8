#!/bin/bash
# round-2 GPU call J: low-latency publication in the triangular solves -- tests, chain trace, timings, full suite
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
echo "== linalg tests" > $O/r2j_tests.log
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "cholesky or paired or large_fact or stepwise or stored" >> $O/r2j_tests.log 2>&1
echo "rc=$?" >> $O/r2j_tests.log
echo "== trace" > $O/r2j_trace.log
timeout 300 python tools/trace_trsm.py 50048 20 0 >> $O/r2j_trace.log 2>&1
timeout 300 python tools/trace_trsm.py 50048 20 1 >> $O/r2j_trace.log 2>&1
echo "== bench_linalg" > $O/r2j_linalg.log
BENCH_M=16384,50048,100096 timeout 600 python tools/gpu_check.py bench_linalg >> $O/r2j_linalg.log 2>&1
for c in 8 32; do
  SAB_TRSM_CHUNK=$c BENCH_M=50048 timeout 600 python tools/gpu_check.py bench_linalg >> $O/r2j_linalg.log 2>&1
done
echo "== all gpu tests" >> $O/r2j_tests.log
( time timeout 1800 python -m pytest tests -q -m gpu ) >> $O/r2j_tests.log 2>&1
echo "rc=$?" >> $O/r2j_tests.log
grep -E "passed|failed|rc=|real" $O/r2j_tests.log; cat $O/r2j_trace.log | head -12; grep "M=" $O/r2j_linalg.log
