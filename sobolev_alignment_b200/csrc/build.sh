#!/bin/bash
# Builds libsobolev_b200.so (sm_100a only) in-tree so that it travels to the GPU box with the snapshot.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${HERE}/../lib"
mkdir -p "${OUT}" "${HERE}/.obj"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC
       --expt-relaxed-constexpr -Xptxas -v ${SAB_NVCC_EXTRA:-})
pids=()
for f in api kmv prep preproc linalg; do
  ( "${NVCC}" "${FLAGS[@]}" -c "${HERE}/${f}.cu" -o "${HERE}/.obj/${f}.o" > "${HERE}/.obj/${f}.log" 2>&1 ) &
  pids+=($!)
done
rc=0
for p in "${pids[@]}"; do wait "$p" || rc=1; done
if [ "$rc" -ne 0 ]; then cat "${HERE}"/.obj/*.log; exit 1; fi
"${NVCC}" -shared -o "${OUT}/libsobolev_b200.so" "${HERE}"/.obj/{api,kmv,prep,preproc,linalg}.o -cudart static
echo "built ${OUT}/libsobolev_b200.so"
