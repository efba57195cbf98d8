// Input preprocessing of the artificial samples on the device: the chain
// /root/reference/sobolev_alignment/sobolev_alignment.py:830-902 (`_memmap_log_processing`,
// `_frobenius_normalisation`) runs on the host in numpy today -- log10(x + 1), StandardScaler, one scalar
// per data source -- i.e. three full passes over a 12 GB array before `KRRApprox.fit` sees it.  Here:
//   sa_colstats        per-column sum / sum of squares of f(x) - pivot (f = identity or log10(x + 1)), float64
//   sa_transform_rows  out = (f(x) - shift) * scale * gain   (fp32, may be in place) + sum of row norms
// Both are HBM-bound streaming kernels (n*p*4 bytes read, + n*p*4 written for the transform).
#include "../../include/sobolev_b200.h"
#include "common.cuh"

namespace sab {
namespace pre {

__device__ __forceinline__ float fwd(float x, int log10p1) { return log10p1 ? log10f(x + 1.0f) : x; }

constexpr int CS_ROWS = 8;  // blockDim = (32, CS_ROWS): 32 x 4 consecutive columns, 8 interleaved rows

// VEC: every thread owns 4 consecutive columns (128-bit loads, 512 B per warp and row); rows are walked 4 at
// a time so that 4 independent loads are in flight per thread
template <bool VEC>
__global__ void __launch_bounds__(32 * CS_ROWS)
colstats_kernel(const float* __restrict__ X, long long n, int p, long long ldx, int log10p1, long long rows_per_block,
                const float* __restrict__ pivot, double* __restrict__ sum, double* __restrict__ sumsq) {
  constexpr int W = VEC ? 4 : 1;
  __shared__ double s1[CS_ROWS][32 * W], s2[CS_ROWS][32 * W];
  const int c = (blockIdx.x * 32 + threadIdx.x) * W;
  const long long r0 = static_cast<long long>(blockIdx.y) * rows_per_block;
  const long long r1 = min(n, r0 + rows_per_block);
  double a[W], b[W];
  float pv[W];
#pragma unroll
  for (int w = 0; w < W; ++w) {
    a[w] = b[w] = 0.0;
    pv[w] = (pivot && c + w < p) ? __ldg(pivot + c + w) : 0.f;
  }
  if (c < p) {
    // Sums of (y - pivot): with the pivot near the column mean, var = E[(y-c)^2] - E[y-c]^2 does not cancel.
    // float partial sums over 64 rows, folded into float64.
    for (long long r = r0 + threadIdx.y; r < r1; r += CS_ROWS * 64) {
      float fa[W], fb[W];
#pragma unroll
      for (int w = 0; w < W; ++w) fa[w] = fb[w] = 0.f;
      const long long rend = min(r1, r + CS_ROWS * 64);
      long long rr = r;
      if (VEC) {
        for (; rr + 3 * CS_ROWS < rend; rr += 4 * CS_ROWS) {
          float4 v[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) v[u] = __ldcs(reinterpret_cast<const float4*>(X + (rr + u * CS_ROWS) * ldx + c));
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float y[4] = {fwd(v[u].x, log10p1) - pv[0], fwd(v[u].y, log10p1) - pv[1],
                                fwd(v[u].z, log10p1) - pv[2], fwd(v[u].w, log10p1) - pv[3]};
#pragma unroll
            for (int w = 0; w < 4; ++w) {
              fa[w] += y[w];
              fb[w] = fmaf(y[w], y[w], fb[w]);
            }
          }
        }
      }
      for (; rr < rend; rr += CS_ROWS) {
#pragma unroll
        for (int w = 0; w < W; ++w)
          if (c + w < p) {
            const float y = fwd(__ldcs(X + rr * ldx + c + w), log10p1) - pv[w];
            fa[w] += y;
            fb[w] = fmaf(y, y, fb[w]);
          }
      }
#pragma unroll
      for (int w = 0; w < W; ++w) {
        a[w] += fa[w];
        b[w] += fb[w];
      }
    }
  }
#pragma unroll
  for (int w = 0; w < W; ++w) {
    s1[threadIdx.y][threadIdx.x * W + w] = a[w];
    s2[threadIdx.y][threadIdx.x * W + w] = b[w];
  }
  __syncthreads();
  if (threadIdx.y == 0) {
#pragma unroll
    for (int w = 0; w < W; ++w) {
      if (c + w >= p) continue;
      double ta = a[w], tb = b[w];
#pragma unroll
      for (int k = 1; k < CS_ROWS; ++k) {
        ta += s1[k][threadIdx.x * W + w];
        tb += s2[k][threadIdx.x * W + w];
      }
      atomicAdd(sum + c + w, ta);
      atomicAdd(sumsq + c + w, tb);
    }
  }
}

constexpr int TR_WARPS = 8;

template <bool VEC>
__global__ void __launch_bounds__(TR_WARPS * 32)
transform_rows_kernel(const float* __restrict__ X, long long n, int p, long long ldx, int log10p1,
                      const float* __restrict__ shift, const float* __restrict__ scale, float gain,
                      float* __restrict__ out, long long ldo, double* __restrict__ rownorm_sum) {
  const int lane = threadIdx.x & 31;
  const long long warp_global = static_cast<long long>(blockIdx.x) * TR_WARPS + (threadIdx.x >> 5);
  const long long warp_stride = static_cast<long long>(gridDim.x) * TR_WARPS;
  double local = 0.0;
  for (long long row = warp_global; row < n; row += warp_stride) {
    const float* x = X + row * ldx;
    float* o = out + row * ldo;
    float acc = 0.f;
    int j0 = 0;
    if (VEC) {
      const int p4 = p >> 2;
      for (int q = lane; q < p4; q += 32) {
        const float4 v = reinterpret_cast<const float4*>(x)[q];
        float4 sh = make_float4(0.f, 0.f, 0.f, 0.f), sc = make_float4(1.f, 1.f, 1.f, 1.f);
        if (shift) sh = __ldg(reinterpret_cast<const float4*>(shift) + q);
        if (scale) sc = __ldg(reinterpret_cast<const float4*>(scale) + q);
        float4 y;
        y.x = (fwd(v.x, log10p1) - sh.x) * sc.x * gain;
        y.y = (fwd(v.y, log10p1) - sh.y) * sc.y * gain;
        y.z = (fwd(v.z, log10p1) - sh.z) * sc.z * gain;
        y.w = (fwd(v.w, log10p1) - sh.w) * sc.w * gain;
        acc = fmaf(y.x, y.x, fmaf(y.y, y.y, fmaf(y.z, y.z, fmaf(y.w, y.w, acc))));
        if (out != nullptr) reinterpret_cast<float4*>(o)[q] = y;
      }
      j0 = p4 << 2;
    }
    for (int j = j0 + lane; j < p; j += 32) {
      const float sh = shift ? __ldg(shift + j) : 0.f;
      const float sc = scale ? __ldg(scale + j) : 1.f;
      const float y = (fwd(x[j], log10p1) - sh) * sc * gain;
      acc = fmaf(y, y, acc);
      if (out != nullptr) o[j] = y;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    local += static_cast<double>(sqrtf(acc));
  }
  if (rownorm_sum != nullptr && lane == 0) atomicAdd(rownorm_sum, local);
}

}  // namespace pre
}  // namespace sab

extern "C" int sa_colstats(void* stream, const float* X, int64_t n, int64_t p, int64_t ldx, int log10p1,
                           const float* pivot, double* sum, double* sumsq) {
  using namespace sab;
  SAB_REQUIRE(n >= 0 && p > 0 && p < (1ll << 30) && ldx >= p, "bad shape n=%lld p=%lld ldx=%lld", (long long)n,
              (long long)p, (long long)ldx);
  SAB_REQUIRE(sum != nullptr && sumsq != nullptr, "NULL output");
  if (n == 0) return 0;
  const long long rows_per_block = 4096;
  const bool vec = (ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0);
  dim3 grid(static_cast<unsigned>(ceil_div(p, vec ? 128 : 32)), static_cast<unsigned>(ceil_div(n, rows_per_block)));
  SAB_REQUIRE(grid.y <= 65535, "too many rows for one call (%lld): call per row chunk", (long long)n);
  if (vec)
    pre::colstats_kernel<true><<<grid, dim3(32, pre::CS_ROWS), 0, static_cast<cudaStream_t>(stream)>>>(
        X, n, static_cast<int>(p), ldx, log10p1, rows_per_block, pivot, sum, sumsq);
  else
    pre::colstats_kernel<false><<<grid, dim3(32, pre::CS_ROWS), 0, static_cast<cudaStream_t>(stream)>>>(
        X, n, static_cast<int>(p), ldx, log10p1, rows_per_block, pivot, sum, sumsq);
  SAB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int sa_transform_rows(void* stream, const float* X, int64_t n, int64_t p, int64_t ldx, int log10p1,
                                 const float* shift, const float* scale, float gain, float* out, int64_t ldo,
                                 double* rownorm_sum) {
  using namespace sab;
  SAB_REQUIRE(n >= 0 && p > 0 && p < (1ll << 30) && ldx >= p && (out == nullptr || ldo >= p), "bad shape n=%lld p=%lld", (long long)n,
              (long long)p);
  SAB_REQUIRE(out != nullptr || rownorm_sum != nullptr, "NULL output");
  if (n == 0) return 0;
  int dev = 0, sms = 0;
  SAB_CHECK_CUDA(cudaGetDevice(&dev));
  SAB_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  long long blocks = ceil_div(n, pre::TR_WARPS);
  if (blocks > static_cast<long long>(sms) * 8) blocks = static_cast<long long>(sms) * 8;
  const bool vec = (ldx % 4 == 0) && (ldo % 4 == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0) &&
                   (out == nullptr || (reinterpret_cast<uintptr_t>(out) & 15) == 0) &&
                   (!shift || (reinterpret_cast<uintptr_t>(shift) & 15) == 0) &&
                   (!scale || (reinterpret_cast<uintptr_t>(scale) & 15) == 0);
  if (vec)
    pre::transform_rows_kernel<true><<<static_cast<int>(blocks), pre::TR_WARPS * 32, 0, static_cast<cudaStream_t>(stream)>>>(
        X, n, static_cast<int>(p), ldx, log10p1, shift, scale, gain, out, ldo, rownorm_sum);
  else
    pre::transform_rows_kernel<false><<<static_cast<int>(blocks), pre::TR_WARPS * 32, 0, static_cast<cudaStream_t>(stream)>>>(
        X, n, static_cast<int>(p), ldx, log10p1, shift, scale, gain, out, ldo, rownorm_sum);
  SAB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
