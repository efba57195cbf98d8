#!/bin/bash
# round-2 GPU call F: stored kernel blocks (design B) -- parity, micro-benchmark, fit; fused-kernel ncu capture
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
echo "== design B tests" > $O/r2f_tests.log
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fit.py -q -m gpu -x -k "stored" >> $O/r2f_tests.log 2>&1
echo "rc=$?" >> $O/r2f_tests.log
timeout 600 python tools/bench_store.py > $O/r2f_store.log 2>&1
echo "== bench with the stored-blocks leg" > $O/r2f_bench.log
timeout 900 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline >> $O/r2f_bench.log 2>&1
echo "rc=$?" >> $O/r2f_bench.log
B="python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-extras"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:kmv_kernel<1, 24, 0" --launch-skip 3 --launch-count 2 -f -o $O/r02_kmv_fused_c3 $B > $O/r2f_ncu_kmv.log 2>&1
echo "rc=$?" >> $O/r2f_ncu_kmv.log
grep -E "passed|failed|rc=" $O/r2f_tests.log; cat $O/r2f_store.log; tail -2 $O/r2f_bench.log | cut -c1-1500; ls -la $O/*.ncu-rep
