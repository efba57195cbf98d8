#!/bin/bash
# Runs each developer check in its own process with a timeout; logs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for what in "$@"; do
  echo "=== $what ===" 
  timeout 300 python tools/gpu_check.py $what > gpurun_out/check_$what.log 2>&1
  echo "exit $?" >> gpurun_out/check_$what.log
  tail -n 45 gpurun_out/check_$what.log
done
