#!/bin/bash
# Builds libsobolev_b200.so (sm_100a only) in-tree so that it travels to the GPU box with the snapshot.
# Objects are rebuilt only when their source (or a shared header) is newer; SAB_REBUILD=1 forces all.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${HERE}/../lib"
mkdir -p "${OUT}" "${HERE}/.obj"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC
       --expt-relaxed-constexpr -Xptxas -v ${SAB_NVCC_EXTRA:-})
SRCS=(api kmv prep preproc linalg cg trsm)
HDRS=("${HERE}/common.cuh" "${HERE}/ptx.cuh" "${HERE}/../../include/sobolev_b200.h" "${HERE}/build.sh")
pids=()
OBJS=()
for f in "${SRCS[@]}"; do
  obj="${HERE}/.obj/${f}.o"
  OBJS+=("${obj}")
  stale=0
  if [ ! -f "${obj}" ] || [ "${HERE}/${f}.cu" -nt "${obj}" ] || [ -n "${SAB_REBUILD:-}" ]; then stale=1; fi
  for h in "${HDRS[@]}"; do if [ "${h}" -nt "${obj}" ]; then stale=1; fi; done
  if [ "${stale}" -eq 1 ]; then
    ( "${NVCC}" "${FLAGS[@]}" -c "${HERE}/${f}.cu" -o "${obj}.tmp" > "${HERE}/.obj/${f}.log" 2>&1 \
        && mv "${obj}.tmp" "${obj}" ) &
    pids+=($!)
  fi
done
rc=0
for p in "${pids[@]:-}"; do [ -z "${p}" ] || wait "$p" || rc=1; done
if [ "$rc" -ne 0 ]; then cat "${HERE}"/.obj/*.log; exit 1; fi
"${NVCC}" -shared -o "${OUT}/libsobolev_b200.so" "${OBJS[@]}" -cudart static
echo "built ${OUT}/libsobolev_b200.so"
