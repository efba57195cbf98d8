#!/bin/bash
# after the chunk-size change: solve tests + C3 bench without the secondary legs
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_scale.py -q -m gpu -k "cholesky or paired or large_fact or matern_many or config2" > $O/r2m_tests.log 2>&1
echo "rc=$?" >> $O/r2m_tests.log
timeout 900 python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > $O/r2m_bench.log 2>&1
echo "rc=$?" >> $O/r2m_bench.log
grep -E "passed|failed|rc=" $O/r2m_tests.log; tail -2 $O/r2m_bench.log | cut -c1-1800
