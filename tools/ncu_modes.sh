#!/bin/bash
# cycles / tensor-pipe activity of the fused kernel under each SAB_DEBUG timing experiment (see kmv.cu)
mkdir -p gpurun_out
M="sm__cycles_elapsed.max,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum,sm__cycles_elapsed.avg.per_second,l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed"
for cfg in "$@"; do
  IFS=: read -r dbg cg <<< "$cfg"
  echo "=== SAB_DEBUG=$dbg SAB_CTA_GROUP=${cg:-2}"
  SAB_DEBUG=$dbg SAB_CTA_GROUP=${cg:-2} timeout 200 ncu --metrics $M --clock-control none -k regex:kmv_kernel -s 1 -c 1 \
     python tools/profile_kmv.py fwd 2>&1 | grep -E "sm__|gpu__|l1tex__"
done
