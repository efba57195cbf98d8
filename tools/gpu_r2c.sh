#!/bin/bash
# round-2 GPU call C: new host features (consensus, null model, grid, fused preprocessing), tuned trsm
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
echo "== new tests first" > $O/r2c_tests.log
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fit.py -q -m gpu -x -k "paired or cholesky or stepwise or fused_pre or consensus or null_model or grid" >> $O/r2c_tests.log 2>&1
echo "rc=$?" >> $O/r2c_tests.log
echo "== trace" > $O/r2c_trace.log
timeout 300 python tools/trace_trsm.py 50048 20 0 >> $O/r2c_trace.log 2>&1
echo "== bench_linalg" > $O/r2c_linalg.log
BENCH_M=50048 timeout 600 python tools/gpu_check.py bench_linalg >> $O/r2c_linalg.log 2>&1
SAB_TRSM_CHUNK=32 BENCH_M=50048 timeout 600 python tools/gpu_check.py bench_linalg >> $O/r2c_linalg.log 2>&1
echo "== all gpu tests (per file)" >> $O/r2c_tests.log
for f in tests/test_gpu_kernels.py tests/test_gpu_fit.py tests/test_gpu_scale.py; do
  timeout 1500 python -m pytest $f -q -m gpu >> $O/r2c_tests.log 2>&1
  echo "rc=$? ($f)" >> $O/r2c_tests.log
done
grep -E "passed|failed|rc=|FAILED" $O/r2c_tests.log; cat $O/r2c_trace.log | head -12; tail -4 $O/r2c_linalg.log
