#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
B="python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-extras"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:trsm_solve_kernel --launch-skip 6 --launch-count 2 -f -o $O/r02_trsm_final_c3 $B > $O/r2n_ncu_trsm.log 2>&1
echo "rc=$?" >> $O/r2n_ncu_trsm.log
timeout 300 python tools/trace_trsm.py 50048 20 0 > $O/r2n_trace.log 2>&1
ls -la $O/r02_trsm_final_c3.ncu-rep; head -9 $O/r2n_trace.log
