// Row preparation: centre, scale, round to fp16 and take the fp32 squared norm OF THE ROUNDED ROW,
// so that |a|^2 + |b|^2 - 2 a.b in the fused kernel is the exact squared distance between the
// rounded points (fp32 accumulation error aside).  HBM-bound: n*p*(4+2) bytes, one warp per row,
// 128-bit loads / 64-bit stores when the row is aligned.
//
// Replaces falkon `square_norm` + the fp32 operand staging of `fmmv` [external dependency],
// reached from /root/reference/sobolev_alignment/krr_approx.py:227,314.
#include <cuda_fp16.h>

#include "../../include/sobolev_b200.h"
#include "common.cuh"

namespace sab {

constexpr int PREP_WARPS = 8;

// LOG: the rows are raw counts and f(x) = log10(x + 1) is applied first -- the reference's preprocessing chain
// (log10(x + 1), StandardScaler, Frobenius gain: sobolev_alignment.py:851-902) is affine after the log, so it
// folds into shift / scale and the fp32 intermediate is never written.
template <bool LOG>
__device__ __forceinline__ float pre_fn(float x) { return LOG ? log10f(x + 1.0f) : x; }

template <bool VEC, bool LOG>
__global__ void __launch_bounds__(PREP_WARPS * 32)
prep_kernel(const float* __restrict__ X, long long n, int p, long long ldx,
            const float* __restrict__ shift, const float* __restrict__ scale, float gscale,
            __half* __restrict__ out, long long ldo, float* __restrict__ norms, long long norms_len) {
  const int lane = threadIdx.x & 31;
  const long long warp_global = static_cast<long long>(blockIdx.x) * PREP_WARPS + (threadIdx.x >> 5);
  const long long warp_stride = static_cast<long long>(gridDim.x) * PREP_WARPS;
  for (long long row = warp_global; row < norms_len; row += warp_stride) {
    if (row >= n) {  // padding of the norm array
      if (lane == 0) norms[row] = 0.f;
      continue;
    }
    const float* x = X + row * ldx;
    __half* o = out + row * ldo;
    float acc = 0.f;
    if (VEC) {
      const int p4 = p >> 2;
      for (int q = lane; q < p4; q += 32) {
        const float4 v = __ldcs(reinterpret_cast<const float4*>(x) + q);
        float4 sh = make_float4(0.f, 0.f, 0.f, 0.f), sc = make_float4(1.f, 1.f, 1.f, 1.f);
        if (shift) sh = __ldg(reinterpret_cast<const float4*>(shift) + q);
        if (scale) sc = __ldg(reinterpret_cast<const float4*>(scale) + q);
        const __half h0 = __float2half_rn((pre_fn<LOG>(v.x) - sh.x) * sc.x * gscale);
        const __half h1 = __float2half_rn((pre_fn<LOG>(v.y) - sh.y) * sc.y * gscale);
        const __half h2 = __float2half_rn((pre_fn<LOG>(v.z) - sh.z) * sc.z * gscale);
        const __half h3 = __float2half_rn((pre_fn<LOG>(v.w) - sh.w) * sc.w * gscale);
        const float f0 = __half2float(h0), f1 = __half2float(h1), f2 = __half2float(h2),
                    f3 = __half2float(h3);
        acc = fmaf(f0, f0, acc);
        acc = fmaf(f1, f1, acc);
        acc = fmaf(f2, f2, acc);
        acc = fmaf(f3, f3, acc);
        __half2 lo = __halves2half2(h0, h1), hi = __halves2half2(h2, h3);
        uint2 pk;
        pk.x = *reinterpret_cast<unsigned int*>(&lo);
        pk.y = *reinterpret_cast<unsigned int*>(&hi);
        reinterpret_cast<uint2*>(o)[q] = pk;
      }
      for (int j = (p4 << 2) + lane; j < p; j += 32) {
        const float sh = shift ? __ldg(shift + j) : 0.f;
        const float sc = scale ? __ldg(scale + j) : 1.f;
        const __half h = __float2half_rn((pre_fn<LOG>(x[j]) - sh) * sc * gscale);
        const float f = __half2float(h);
        acc = fmaf(f, f, acc);
        o[j] = h;
      }
    } else {
      for (int j = lane; j < p; j += 32) {
        const float sh = shift ? __ldg(shift + j) : 0.f;
        const float sc = scale ? __ldg(scale + j) : 1.f;
        const __half h = __float2half_rn((pre_fn<LOG>(x[j]) - sh) * sc * gscale);
        const float f = __half2float(h);
        acc = fmaf(f, f, acc);
        o[j] = h;
      }
    }
    for (long long j = p + lane; j < ldo; j += 32) o[j] = __float2half_rn(0.f);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) norms[row] = acc;
  }
}

}  // namespace sab

namespace sab {
static int prep_launch(void* stream, const float* X, int64_t n, int64_t p, int64_t ldx, const float* shift,
                       const float* scale, float gscale, void* out_f16, int64_t ldo, float* norms, int64_t norms_len,
                       bool log10p1) {
  SAB_REQUIRE(n >= 0 && p > 0 && p < (1ll << 30), "bad shape n=%lld p=%lld", (long long)n, (long long)p);
  SAB_REQUIRE(ldx >= p && ldo >= p && ldo % 64 == 0, "need ldx >= p, ldo >= p and ldo %% 64 == 0 (ldx=%lld ldo=%lld)",
              (long long)ldx, (long long)ldo);
  SAB_REQUIRE(norms_len >= n, "norms_len (%lld) < n (%lld)", (long long)norms_len, (long long)n);
  if (norms_len == 0) return 0;
  const bool vec = (ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0) &&
                   (!shift || (reinterpret_cast<uintptr_t>(shift) & 15) == 0) &&
                   (!scale || (reinterpret_cast<uintptr_t>(scale) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(out_f16) & 7) == 0);
  int dev = 0, sms = 0;
  SAB_CHECK_CUDA(cudaGetDevice(&dev));
  SAB_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  long long blocks = ceil_div(norms_len, PREP_WARPS);
  const long long max_blocks = static_cast<long long>(sms) * 8;
  if (blocks > max_blocks) blocks = max_blocks;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  auto go = [&](auto kern) {
    kern<<<static_cast<int>(blocks), PREP_WARPS * 32, 0, st>>>(X, n, static_cast<int>(p), ldx, shift, scale, gscale,
                                                               static_cast<__half*>(out_f16), ldo, norms, norms_len);
  };
  if (vec && log10p1) go(prep_kernel<true, true>);
  else if (vec) go(prep_kernel<true, false>);
  else if (log10p1) go(prep_kernel<false, true>);
  else go(prep_kernel<false, false>);
  SAB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
}  // namespace sab

extern "C" int sa_prep(void* stream, const float* X, int64_t n, int64_t p, int64_t ldx,
                       const float* shift, const float* scale, float gscale, void* out_f16,
                       int64_t ldo, float* norms, int64_t norms_len) {
  return sab::prep_launch(stream, X, n, p, ldx, shift, scale, gscale, out_f16, ldo, norms, norms_len, false);
}

extern "C" int sa_prep_log(void* stream, const float* X, int64_t n, int64_t p, int64_t ldx,
                           const float* shift, const float* scale, float gscale, void* out_f16,
                           int64_t ldo, float* norms, int64_t norms_len) {
  return sab::prep_launch(stream, X, n, p, ldx, shift, scale, gscale, out_f16, ldo, norms, norms_len, true);
}
