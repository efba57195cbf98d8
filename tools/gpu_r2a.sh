#!/bin/bash
# round-2 GPU call A: new triangular solves / CG kernels, parity suites, first bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/r2a_smi.log 2>&1
echo "== linalg + cg kernel tests" > $O/r2a_tests.log
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "cholesky or paired or large_fact or cg_vector or ltl_parts" >> $O/r2a_tests.log 2>&1
echo "rc=$?" >> $O/r2a_tests.log
echo "== bench_linalg" > $O/r2a_linalg.log
BENCH_M=16384,50048 timeout 600 python tools/gpu_check.py bench_linalg >> $O/r2a_linalg.log 2>&1
for c in 8 32 64; do
  echo "-- SAB_TRSM_CHUNK=$c" >> $O/r2a_linalg.log
  SAB_TRSM_CHUNK=$c BENCH_M=50048 timeout 600 python tools/gpu_check.py bench_linalg >> $O/r2a_linalg.log 2>&1
done
echo "== all gpu tests (per file)" >> $O/r2a_tests.log
for f in tests/test_gpu_kernels.py tests/test_gpu_fit.py tests/test_gpu_scale.py; do
  timeout 1500 python -m pytest $f -q -m gpu -s >> $O/r2a_tests.log 2>&1
  echo "rc=$? ($f)" >> $O/r2a_tests.log
done
echo "== bench C3" > $O/r2a_bench.log
timeout 900 python bench.py --steps 3 --warmup 1 >> $O/r2a_bench.log 2>&1
echo "rc=$?" >> $O/r2a_bench.log
tail -3 $O/r2a_tests.log; tail -5 $O/r2a_linalg.log; tail -2 $O/r2a_bench.log
