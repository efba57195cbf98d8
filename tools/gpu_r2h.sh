#!/bin/bash
# round-2 GPU call H: full GPU suites, stored-blocks micro-benchmark, fused-kernel ncu capture, default bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
echo "== all gpu tests" > $O/r2h_tests.log
( time timeout 1800 python -m pytest tests -q -m gpu ) >> $O/r2h_tests.log 2>&1
echo "rc=$?" >> $O/r2h_tests.log
timeout 600 python tools/bench_store.py > $O/r2h_store.log 2>&1
B="python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-extras"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:kmv_kernel<\(int\)1, \(int\)24, \(int\)0' --launch-skip 3 --launch-count 2 -f -o $O/r02_kmv_fused_c3 $B > $O/r2h_ncu_kmv.log 2>&1
echo "rc=$?" >> $O/r2h_ncu_kmv.log
echo "== bench default" > $O/r2h_bench.log
( time timeout 1500 python bench.py ) >> $O/r2h_bench.log 2>&1
echo "rc=$?" >> $O/r2h_bench.log
python __graft_entry__.py smoke > $O/r2h_smoke.log 2>&1
grep -E "passed|failed|rc=|real" $O/r2h_tests.log; cat $O/r2h_store.log | head -3; tail -2 $O/r2h_ncu_kmv.log; tail -3 $O/r2h_smoke.log; ls -la $O/*.ncu-rep
