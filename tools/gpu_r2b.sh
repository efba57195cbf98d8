#!/bin/bash
# round-2 GPU call B: tuned triangular solve (+trace), stepwise Cholesky, full test suites
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
echo "== linalg tests" > $O/r2b_tests.log
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "cholesky or paired or large_fact or cg_vector or ltl_parts or stepwise" >> $O/r2b_tests.log 2>&1
echo "rc=$?" >> $O/r2b_tests.log
echo "== trace" > $O/r2b_trace.log
timeout 300 python tools/trace_trsm.py 50048 20 0 >> $O/r2b_trace.log 2>&1
timeout 300 python tools/trace_trsm.py 50048 20 1 >> $O/r2b_trace.log 2>&1
echo "== bench_linalg" > $O/r2b_linalg.log
BENCH_M=16384,50048 timeout 600 python tools/gpu_check.py bench_linalg >> $O/r2b_linalg.log 2>&1
for c in 8 32; do
  echo "-- SAB_TRSM_CHUNK=$c" >> $O/r2b_linalg.log
  SAB_TRSM_CHUNK=$c BENCH_M=50048 timeout 600 python tools/gpu_check.py bench_linalg >> $O/r2b_linalg.log 2>&1
done
echo "== all gpu tests (per file)" >> $O/r2b_tests.log
for f in tests/test_gpu_kernels.py tests/test_gpu_fit.py tests/test_gpu_scale.py; do
  timeout 1500 python -m pytest $f -q -m gpu >> $O/r2b_tests.log 2>&1
  echo "rc=$? ($f)" >> $O/r2b_tests.log
done
grep -E "passed|failed|rc=" $O/r2b_tests.log; cat $O/r2b_trace.log; tail -8 $O/r2b_linalg.log
