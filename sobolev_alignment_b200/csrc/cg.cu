// Conjugate-gradient vector kernels over row-major M x dp fp32 arrays (dp <= 32).
//
// Replaces falkon `ConjugateGradient.solve` [external dependency], reached from
// /root/reference/sobolev_alignment/krr_approx.py:227.
//
// Every column reduction is DETERMINISTIC: a block folds its rows in a fixed order into one partial
// per column, stores it in its own slot of a scratch array, and the block that finishes last sums the
// slots in index order.  Ranks that hold replicated CG state therefore stay bit-identical (no float
// atomics anywhere), which the row-sharded solver relies on: every rank takes the same convergence
// decision without a broadcast.
//
// Convergence is decided ON THE DEVICE: the finishing block writes the column sums to `hist` (one row
// of the residual history) and raises `*stop` when max_c sqrt(rs[c]) < tol.  Once raised, sa_cg_update
// and sa_cg_direction leave X, R, P and the scalars untouched, so the host may read the flag late (an
// asynchronous copy, no per-iteration synchronisation) without changing the result.
#include <cstdlib>

#include "../../include/sobolev_b200.h"
#include "common.cuh"

namespace sab {
namespace cg {

constexpr int RED_THREADS = 512;
constexpr int MAX_GRID = 592;          // 4 blocks per SM
constexpr int WS_COUNTER = MAX_GRID * 32;   // float index of the arrival counter inside the scratch

// blockDim = (dp, rows): thread (c, ry) walks rows ry, ry + stride, ... of column c, so a warp reads
// consecutive floats.  Returns true in the threads of the block that arrived last, with the final sums
// in sums[c] (shared).
__device__ __forceinline__ bool column_reduce(float v, int dp, float* __restrict__ ws, float* sums) {
  __shared__ float red[RED_THREADS];
  __shared__ bool last;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  red[tid] = v;
  __syncthreads();
  if (threadIdx.y == 0) {
    float s = 0.f;
    for (int r = 0; r < static_cast<int>(blockDim.y); ++r) s += red[r * dp + threadIdx.x];
    __stcg(ws + static_cast<long long>(blockIdx.x) * 32 + threadIdx.x, s);
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    unsigned* counter = reinterpret_cast<unsigned*>(ws + WS_COUNTER);
    const unsigned seen = atomicAdd(counter, 1u);
    last = seen + 1 == gridDim.x;
    if (last) *counter = 0;   // ready for the next launch on this stream
  }
  __syncthreads();
  if (!last) return false;
  __threadfence();
  if (threadIdx.y == 0) {
    float s = 0.f;
    for (unsigned b = 0; b < gridDim.x; ++b) s += __ldcg(ws + static_cast<long long>(b) * 32 + threadIdx.x);
    sums[threadIdx.x] = s;
  }
  __syncthreads();
  return true;
}

// residual bookkeeping of the finishing block: history row + stop flag
__device__ __forceinline__ void publish_residual(const float* sums, int dp, float* __restrict__ hist,
                                                 int* __restrict__ stop, float tol) {
  if (threadIdx.y != 0) return;
  if (hist != nullptr) hist[threadIdx.x] = sums[threadIdx.x];
  if (stop != nullptr && threadIdx.x == 0) {
    float m = 0.f;
    for (int c = 0; c < dp; ++c) m = fmaxf(m, sums[c]);
    if (sqrtf(m) < tol) *stop = 1;
  }
}

__global__ void axpby_kernel(long long len, float a, const float* __restrict__ X, float b,
                             float* __restrict__ Y) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < len; i += stride)
    Y[i] = (b == 0.f) ? a * X[i] : fmaf(a, X[i], b * Y[i]);
}

__global__ void coldot_kernel(const float* __restrict__ P, const float* __restrict__ Q, long long M, int dp,
                              float* __restrict__ out, float* __restrict__ ws, float* __restrict__ hist,
                              int* __restrict__ stop, float tol) {
  __shared__ float sums[32];
  float s = 0.f;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.y;
  for (long long r = static_cast<long long>(blockIdx.x) * blockDim.y + threadIdx.y; r < M; r += stride)
    s = fmaf(P[r * dp + threadIdx.x], Q[r * dp + threadIdx.x], s);
  if (!column_reduce(s, dp, ws, sums)) return;
  if (threadIdx.y == 0) out[threadIdx.x] = sums[threadIdx.x];
  publish_residual(sums, dp, hist, stop, tol);
}

__global__ void cg_update_kernel(long long M, int dp, const float* __restrict__ P, const float* __restrict__ AP,
                                 float* __restrict__ X, float* __restrict__ R, const float* __restrict__ rs_old,
                                 const float* __restrict__ pAp, float eps, float* __restrict__ rs_new,
                                 float* __restrict__ ws, float* __restrict__ hist, int* __restrict__ stop,
                                 float tol) {
  __shared__ float sums[32];
  if (stop != nullptr && *reinterpret_cast<volatile int*>(stop) != 0) return;   // converged earlier: frozen
  const int c = threadIdx.x;
  const float alpha = rs_old[c] / (pAp[c] + eps);
  float s = 0.f;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.y;
  for (long long r = static_cast<long long>(blockIdx.x) * blockDim.y + threadIdx.y; r < M; r += stride) {
    const long long i = r * dp + c;
    X[i] = fmaf(alpha, P[i], X[i]);
    if (AP != nullptr) {
      const float rv = fmaf(-alpha, AP[i], R[i]);
      R[i] = rv;
      s = fmaf(rv, rv, s);
    }
  }
  if (AP == nullptr) return;
  if (!column_reduce(s, dp, ws, sums)) return;
  if (threadIdx.y == 0) rs_new[c] = sums[c];
  publish_residual(sums, dp, hist, stop, tol);
}

// P = R + rs_new / (rs_old + eps) P   (falkon divides by Rsold + cg_epsilon)
__global__ void cg_direction_kernel(long long M, int dp, const float* __restrict__ R, float* __restrict__ P,
                                    const float* __restrict__ rs_new, const float* __restrict__ rs_old, float eps,
                                    const int* __restrict__ stop) {
  if (stop != nullptr && *reinterpret_cast<const volatile int*>(stop) != 0) return;
  const int c = threadIdx.x;
  const float ro = rs_old[c];
  const float beta = ro > 0.f ? rs_new[c] / (ro + eps) : 0.f;  // padded / exactly converged columns
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.y;
  for (long long r = static_cast<long long>(blockIdx.x) * blockDim.y + threadIdx.y; r < M; r += stride) {
    const long long i = r * dp + c;
    P[i] = fmaf(beta, P[i], R[i]);
  }
}

static int red_grid(long long M, int rows_per_block) {
  long long g = ceil_div(M, static_cast<long long>(rows_per_block) * 4);
  if (g < 1) g = 1;
  if (g > MAX_GRID) g = MAX_GRID;
  return static_cast<int>(g);
}

}  // namespace cg
}  // namespace sab

using namespace sab;
using namespace sab::cg;

#define SAB_REQUIRE_VEC(M, dp) \
  SAB_REQUIRE((M) > 0 && (dp) >= 1 && (dp) <= 32, "bad vector shape M=%lld dp=%d", (long long)(M), (dp))
#define SAB_REQUIRE_CGWS(ws) \
  SAB_REQUIRE((ws) != nullptr && (reinterpret_cast<uintptr_t>(ws) & 15) == 0, "CG scratch must be 16-byte aligned")

extern "C" int64_t sa_cg_workspace_bytes(void) { return (WS_COUNTER + 32) * 4; }

extern "C" int sa_coldot(void* stream, const float* P, const float* Q, int64_t M, int dp, float* out,
                         void* workspace, float* hist, int* stop, float tol) {
  SAB_REQUIRE_VEC(M, dp);
  SAB_REQUIRE_CGWS(workspace);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int rows = RED_THREADS / dp;
  coldot_kernel<<<red_grid(M, rows), dim3(dp, rows), 0, st>>>(P, Q, M, dp, out, static_cast<float*>(workspace),
                                                               hist, stop, tol);
  SAB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int sa_cg_update(void* stream, int64_t M, int dp, const float* P, const float* AP, float* X,
                            float* R, const float* rs_old, const float* pAp, float eps, float* rs_new,
                            void* workspace, float* hist, int* stop, float tol) {
  SAB_REQUIRE_VEC(M, dp);
  SAB_REQUIRE_CGWS(workspace);
  SAB_REQUIRE(AP == nullptr || (R != nullptr && rs_new != nullptr), "R / rs_new are required with AP");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int rows = RED_THREADS / dp;
  cg_update_kernel<<<red_grid(M, rows), dim3(dp, rows), 0, st>>>(M, dp, P, AP, X, R, rs_old, pAp, eps, rs_new,
                                                                  static_cast<float*>(workspace), hist, stop, tol);
  SAB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int sa_cg_direction(void* stream, int64_t M, int dp, const float* R, float* P, const float* rs_new,
                               const float* rs_old, float eps, const int* stop) {
  SAB_REQUIRE_VEC(M, dp);
  const int rows = RED_THREADS / dp;
  cg_direction_kernel<<<red_grid(M, rows), dim3(dp, rows), 0, static_cast<cudaStream_t>(stream)>>>(
      M, dp, R, P, rs_new, rs_old, eps, stop);
  SAB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int sa_axpby(void* stream, int64_t len, float a, const float* X, float b, float* Y) {
  SAB_REQUIRE(len >= 0, "negative length");
  if (len == 0) return 0;
  long long g = ceil_div(len, 256 * 4);
  if (g > 1184) g = 1184;
  axpby_kernel<<<static_cast<int>(g), 256, 0, static_cast<cudaStream_t>(stream)>>>(len, a, X, b, Y);
  SAB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
