#!/bin/bash
# round-2 GPU call I: fused-kernel ncu capture (fixed kernel filter) + default bench with the grid leg
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
B="python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-extras"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:.*kmv_kernel<\(int\)1, \(int\)20, \(int\)0.*' --launch-skip 3 --launch-count 2 -f -o $O/r02_kmv_fused_c3 $B > $O/r2i_ncu_kmv.log 2>&1
echo "rc=$?" >> $O/r2i_ncu_kmv.log
echo "== bench default" > $O/r2i_bench.log
( time timeout 1500 python bench.py ) >> $O/r2i_bench.log 2>&1
echo "rc=$?" >> $O/r2i_bench.log
tail -3 $O/r2i_ncu_kmv.log | cut -c1-300; ls -la $O/*.ncu-rep
