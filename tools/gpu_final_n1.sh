#!/bin/bash
# final single-GPU validation: smoke, the whole GPU suite, the default bench line
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
python __graft_entry__.py smoke > $O/final_smoke.log 2>&1
echo "rc=$?" >> $O/final_smoke.log
( time timeout 1800 python -m pytest tests -x -q -m gpu ) > $O/final_tests.log 2>&1
echo "rc=$?" >> $O/final_tests.log
( time timeout 1500 python bench.py ) > $O/final_bench_n1.log 2>&1
echo "rc=$?" >> $O/final_bench_n1.log
tail -2 $O/final_smoke.log; grep -E "passed|failed|rc=|real" $O/final_tests.log; grep -E "real|rc=" $O/final_bench_n1.log
