#!/bin/bash
# round-2 GPU call D: full default bench (all legs) + ncu evidence (launch list of one C3 fit, full-set captures)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
echo "== bench default" > $O/r2d_bench.log
( time timeout 1500 python bench.py ) >> $O/r2d_bench.log 2>&1
echo "rc=$?" >> $O/r2d_bench.log
echo "== reference arm" >> $O/r2d_bench.log
( time timeout 600 python bench.py --impl reference --steps 1 --warmup 1 ) >> $O/r2d_bench.log 2>&1
B="python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-extras"
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 7000 --csv --log-file $O/r02_c3_fit_launches.csv $B > $O/r2d_ncu_list.log 2>&1
echo "rc=$?" >> $O/r2d_ncu_list.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kmv_kernel --launch-skip 4 --launch-count 2 -f -o $O/r02_kmv_c3 $B > $O/r2d_ncu_kmv.log 2>&1
echo "rc=$?" >> $O/r2d_ncu_kmv.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:trsm_solve_kernel --launch-skip 6 --launch-count 2 -f -o $O/r02_trsm_c3 $B > $O/r2d_ncu_trsm.log 2>&1
echo "rc=$?" >> $O/r2d_ncu_trsm.log
ls -la $O/*.ncu-rep; tail -3 $O/r2d_bench.log | cut -c1-600
