// fp32 dense linear algebra of the Nystrom preconditioner (factorisations).
//
// Everything here works on matrices whose order is a multiple of 128 (the host pads K_MM with an
// identity block, which leaves the Cholesky factors and every solve unchanged), row-major, lower
// triangle significant:
//   sa_potrf   blocked right-looking Cholesky: 128x128 diagonal factor+inverse in shared memory,
//              panel = A * inv(L_kk)^T (SIMT SGEMM), trailing update C -= P P^T on the tcgen05 tensor
//              path: the fp32 panel is split into fp16 pairs x 2^s = hi + lo and multiplied as the
//              K-concatenated product [lo|hi|hi] [hi|lo|hi]^T (lo lo dropped, ~2^-22 relative) by the
//              GEMM mode of the fused kernel (kmv.cu), which keeps the update HBM-bound
//   sa_ltl     G = alpha L^T L + beta I as rank-128 updates with the block rows of L (same split,
//              transposed), also on the tensor path
// The triangular solves live in trsm.cu, the CG vector kernels in cg.cu.
//
// Replaces falkon `FalkonPreconditioner` (potrf, lauum, mul_triang)
// [external dependency], reached from
// /root/reference/sobolev_alignment/krr_approx.py:227.
#include <cuda_fp16.h>

#include <cstdlib>

#include "../../include/sobolev_b200.h"
#include "common.cuh"

namespace sab {
int tc_gemm_nt(void* stream, const void* A, int64_t ra, int64_t lda, const void* B, int64_t rb, int64_t ldb,
               int64_t k, float* out, int64_t ldo, float alpha, const float* alpha_dev, float gamma, int lower);
namespace la {

constexpr int TS = 128;        // tile size
constexpr int KC = 16;         // k chunk of the SGEMM
constexpr int LDT = TS + 4;    // padded shared-memory row of the SGEMM operand tiles

enum { MODE_PANEL = 0, MODE_SYRK = 1, MODE_LTL = 2 };

// --------------------------------------------------------------------------------------------
// 128x128 SIMT SGEMM tile: 256 threads, each 8x8 outputs split as 2x2 groups of 4x4 so that every
// shared-memory read is a conflict-free 128-bit load.
// --------------------------------------------------------------------------------------------
struct GemmSmem {
  float a[2][KC][LDT];
  float b[2][KC][LDT];
};

// operand stored [row][k] in global (k contiguous): transpose while staging
__device__ __forceinline__ void load_nt(const float* __restrict__ src, long long ld, int kc0,
                                        float4 (&reg)[2]) {
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int idx = threadIdx.x + i * 256;
    const int r = idx >> 2, kq = idx & 3;
    reg[i] = *reinterpret_cast<const float4*>(src + static_cast<long long>(r) * ld + kc0 + kq * 4);
  }
}
__device__ __forceinline__ void store_nt(float (*dst)[LDT], const float4 (&reg)[2]) {
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int idx = threadIdx.x + i * 256;
    const int r = idx >> 2, kq = idx & 3;
    dst[kq * 4 + 0][r] = reg[i].x;
    dst[kq * 4 + 1][r] = reg[i].y;
    dst[kq * 4 + 2][r] = reg[i].z;
    dst[kq * 4 + 3][r] = reg[i].w;
  }
}
// operand stored [k][col] in global (col contiguous): straight copy
__device__ __forceinline__ void load_tn(const float* __restrict__ src, long long ld, int kc0,
                                        float4 (&reg)[2]) {
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int idx = threadIdx.x + i * 256;
    const int kk = idx >> 5, cq = idx & 31;
    reg[i] = *reinterpret_cast<const float4*>(src + static_cast<long long>(kc0 + kk) * ld + cq * 4);
  }
}
__device__ __forceinline__ void store_tn(float (*dst)[LDT], const float4 (&reg)[2]) {
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int idx = threadIdx.x + i * 256;
    const int kk = idx >> 5, cq = idx & 31;
    *reinterpret_cast<float4*>(&dst[kk][cq * 4]) = reg[i];
  }
}

template <bool TN>
__device__ __forceinline__ void gemm_tile(const float* __restrict__ A, const float* __restrict__ B,
                                          long long lda, long long ldb, int ksteps, GemmSmem& sm,
                                          float (&acc)[8][8]) {
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  float4 ra[2], rb[2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  if (TN) { load_tn(A, lda, 0, ra); load_tn(B, ldb, 0, rb); store_tn(sm.a[0], ra); store_tn(sm.b[0], rb); }
  else    { load_nt(A, lda, 0, ra); load_nt(B, ldb, 0, rb); store_nt(sm.a[0], ra); store_nt(sm.b[0], rb); }
  __syncthreads();
  int cur = 0;
  for (int kt = 0; kt < ksteps; ++kt) {
    const bool more = kt + 1 < ksteps;
    if (more) {
      if (TN) { load_tn(A, lda, (kt + 1) * KC, ra); load_tn(B, ldb, (kt + 1) * KC, rb); }
      else    { load_nt(A, lda, (kt + 1) * KC, ra); load_nt(B, ldb, (kt + 1) * KC, rb); }
    }
#pragma unroll
    for (int kk = 0; kk < KC; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&sm.a[cur][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&sm.a[cur][kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&sm.b[cur][kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&sm.b[cur][kk][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (more) {
      if (TN) { store_tn(sm.a[cur ^ 1], ra); store_tn(sm.b[cur ^ 1], rb); }
      else    { store_nt(sm.a[cur ^ 1], ra); store_nt(sm.b[cur ^ 1], rb); }
    }
    __syncthreads();
    cur ^= 1;
  }
}

// rows/cols owned by this thread inside the 128x128 tile
__device__ __forceinline__ int tile_row(int i) { return (i < 4 ? 0 : 64) + (threadIdx.x >> 4) * 4 + (i & 3); }
__device__ __forceinline__ int tile_col(int j) { return (j < 4 ? 0 : 64) + (threadIdx.x & 15) * 4; }

__device__ __forceinline__ void tri_index(int t, int& i, int& j) {
  // t = i (i + 1) / 2 + j, 0 <= j <= i
  i = static_cast<int>((sqrtf(8.f * static_cast<float>(t) + 1.f) - 1.f) * 0.5f);
  while ((i + 1) * (i + 2) / 2 <= t) ++i;
  while (i * (i + 1) / 2 > t) --i;
  j = t - i * (i + 1) / 2;
}

// PANEL : A[k0+128+bi*128 .., k0..k0+128) <- A_panel * invL^T          (in place)
// SYRK  : C(i, j) -= L(i, k) L(j, k)^T for block indices i >= j > k
// LTL   : G(i, j)  = alpha * sum_{l >= i} L(l, i)^T L(l, j) + beta * I
template <int MODE>
__global__ void __launch_bounds__(256)
sgemm_kernel(float* __restrict__ A, long long lda, const float* __restrict__ aux, long long ldaux,
             int k, int nblk, float alpha, float beta) {
  __shared__ GemmSmem sm;
  float acc[8][8];
  if (MODE == MODE_PANEL) {
    const long long r0 = static_cast<long long>(k + 1 + blockIdx.x) * TS;
    float* panel = A + r0 * lda + static_cast<long long>(k) * TS;
    gemm_tile<false>(panel, aux, lda, TS, TS / KC, sm, acc);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float* dst = panel + static_cast<long long>(tile_row(i)) * lda;
#pragma unroll
      for (int jh = 0; jh < 2; ++jh)
        *reinterpret_cast<float4*>(dst + tile_col(jh * 4)) =
            make_float4(acc[i][jh * 4], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]);
    }
  } else if (MODE == MODE_SYRK) {
    // jlim = number of block columns right of k that are updated (all of them for the classic
    // right-looking step, the rest of the current outer panel for the two-level factorisation)
    const int jlim = static_cast<int>(alpha);
    int ti, tj;
    if (jlim >= nblk - k - 1) {
      tri_index(blockIdx.x, ti, tj);
    } else {
      ti = static_cast<int>(blockIdx.x) / jlim;
      tj = static_cast<int>(blockIdx.x) % jlim;
      if (tj > ti) return;
    }
    const long long bi = k + 1 + ti, bj = k + 1 + tj;
    const float* Li = A + bi * TS * lda + static_cast<long long>(k) * TS;
    const float* Lj = A + bj * TS * lda + static_cast<long long>(k) * TS;
    gemm_tile<false>(Li, Lj, lda, lda, TS / KC, sm, acc);
    float* C = A + bi * TS * lda + bj * TS;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float* dst = C + static_cast<long long>(tile_row(i)) * lda;
#pragma unroll
      for (int jh = 0; jh < 2; ++jh) {
        float4* p = reinterpret_cast<float4*>(dst + tile_col(jh * 4));
        float4 c = *p;
        c.x -= acc[i][jh * 4];
        c.y -= acc[i][jh * 4 + 1];
        c.z -= acc[i][jh * 4 + 2];
        c.w -= acc[i][jh * 4 + 3];
        *p = c;
      }
    }
  } else {
    int ti, tj;
    tri_index(blockIdx.x, ti, tj);
    const float* L = aux;
    const float* Li = L + static_cast<long long>(ti) * TS * ldaux + static_cast<long long>(ti) * TS;
    const float* Lj = L + static_cast<long long>(ti) * TS * ldaux + static_cast<long long>(tj) * TS;
    gemm_tile<true>(Li, Lj, ldaux, ldaux, (nblk - ti) * (TS / KC), sm, acc);
    float* G = A + static_cast<long long>(ti) * TS * lda + static_cast<long long>(tj) * TS;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = tile_row(i);
      float* dst = G + static_cast<long long>(r) * lda;
#pragma unroll
      for (int jh = 0; jh < 2; ++jh) {
        const int c = tile_col(jh * 4);
        float4 v = make_float4(alpha * acc[i][jh * 4], alpha * acc[i][jh * 4 + 1],
                               alpha * acc[i][jh * 4 + 2], alpha * acc[i][jh * 4 + 3]);
        if (ti == tj) {
          if (c + 0 == r) v.x += beta;
          if (c + 1 == r) v.y += beta;
          if (c + 2 == r) v.z += beta;
          if (c + 3 == r) v.w += beta;
        }
        *reinterpret_cast<float4*>(dst + c) = v;
      }
    }
  }
}

// --------------------------------------------------------------------------------------------
// 128x128 diagonal block: Cholesky, then the explicit inverse of the factor, both register resident.
// 256 threads as a 16 x 16 grid; thread (ty, tx) owns the elements (ty + 16 a, tx + 16 b), a, b < 8
// (cyclic, so the shrinking active region stays balanced).  Every elimination step broadcasts one
// column (and, for the inverse, one row) through shared memory and is a rank-1 update of 64 registers.
// --------------------------------------------------------------------------------------------
constexpr int LDD = TS + 1;

__global__ void __launch_bounds__(256)
diag_factor_kernel(float* __restrict__ A, long long lda, int k, float* __restrict__ invdiag,
                   int* __restrict__ info) {
  extern __shared__ float dsm[];
  float* Ls = dsm;                      // [128][129] staging of the block / of L
  float* colb = dsm + TS * LDD;         // [2][128] column broadcast (double buffered)
  float* rowb = colb + 2 * TS;          // [2][128] row broadcast
  float* blk = A + static_cast<long long>(k) * TS * lda + static_cast<long long>(k) * TS;
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  for (int idx = tid; idx < TS * TS / 4; idx += 256) {
    const int r = idx >> 5, cq = idx & 31;
    const float4 v = *reinterpret_cast<const float4*>(blk + static_cast<long long>(r) * lda + cq * 4);
    Ls[r * LDD + cq * 4 + 0] = v.x;
    Ls[r * LDD + cq * 4 + 1] = v.y;
    Ls[r * LDD + cq * 4 + 2] = v.z;
    Ls[r * LDD + cq * 4 + 3] = v.w;
  }
  __syncthreads();
  float a[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) a[i][j] = Ls[(ty + 16 * i) * LDD + tx + 16 * j];

  // ---- right-looking Cholesky: after step j, column j of `a` holds L[:, j].  The column loop is split
  // as j = 16 jb + jx with jb unrolled, so that every register index and most ownership tests are
  // compile-time (a runtime `b == j / 16` costs an indexed branch per step).
#pragma unroll
  for (int jb = 0; jb < 8; ++jb) {
#pragma unroll 1
    for (int jx = 0; jx < 16; ++jx) {
      const int j = jb * 16 + jx;
      float* cb = colb + (j & 1) * TS;
      if (tx == jx) {
#pragma unroll
        for (int i = jb; i < 8; ++i) cb[ty + 16 * i] = a[i][jb];  // rows >= 16 jb: all that is needed
      }
      __syncthreads();
      const float piv = cb[j];
      if (tid == 0 && !(piv > 0.f)) atomicCAS(info, 0, k * TS + j + 1);
      const float inv = rsqrtf(fmaxf(piv, 1e-30f));
      float li[8], lc[8];
#pragma unroll
      for (int i = jb; i < 8; ++i) li[i] = (i > jb || ty > jx) ? cb[ty + 16 * i] * inv : 0.f;
#pragma unroll
      for (int b = jb; b < 8; ++b) lc[b] = (b > jb || tx > jx) ? cb[tx + 16 * b] * inv : 0.f;
#pragma unroll
      for (int i = jb; i < 8; ++i)
#pragma unroll
        for (int b = jb; b < 8; ++b) a[i][b] = fmaf(-li[i], lc[b], a[i][b]);
      if (tx == jx) {  // column j itself becomes L[:, j]
#pragma unroll
        for (int i = jb; i < 8; ++i) {
          if (i > jb || ty > jx) a[i][jb] = li[i];
          else if (ty == jx) a[i][jb] = piv * inv;  // sqrt(piv)
        }
      }
    }
  }
  __syncthreads();
  // L (lower, zero above the diagonal) back to shared memory
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      const int r = ty + 16 * i, c = tx + 16 * b;
      Ls[r * LDD + c] = c <= r ? a[i][b] : 0.f;
    }
  __syncthreads();
  for (int idx = tid; idx < TS * TS; idx += 256) {
    const int r = idx >> 7, c = idx & 127;
    blk[static_cast<long long>(r) * lda + c] = Ls[r * LDD + c];
  }

  // ---- X = L^-1 by forward elimination on the identity: after step j row j of X is final and
  // X[i, :] -= L[i, j] X[j, :] for i > j.  `a` is reused for X.
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int b = 0; b < 8; ++b) a[i][b] = (ty + 16 * i == tx + 16 * b) ? 1.f : 0.f;
#pragma unroll
  for (int jb = 0; jb < 8; ++jb) {
#pragma unroll 1
    for (int jy = 0; jy < 16; ++jy) {
      const int j = jb * 16 + jy;
      float* rb = rowb + (j & 1) * TS;
      const float dinv = 1.f / Ls[j * LDD + j];
      if (ty == jy) {  // finish row j and broadcast it (X is lower triangular: columns <= j)
#pragma unroll
        for (int b = 0; b <= jb; ++b) {
          a[jb][b] *= dinv;
          rb[tx + 16 * b] = a[jb][b];
        }
      }
      __syncthreads();
      float li[8], xr[8];
#pragma unroll
      for (int i = jb; i < 8; ++i) li[i] = (i > jb || ty > jy) ? Ls[(ty + 16 * i) * LDD + j] : 0.f;
#pragma unroll
      for (int b = 0; b <= jb; ++b) xr[b] = rb[tx + 16 * b];
#pragma unroll
      for (int i = jb; i < 8; ++i)
#pragma unroll
        for (int b = 0; b <= jb; ++b) a[i][b] = fmaf(-li[i], xr[b], a[i][b]);
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int b = 0; b < 8; ++b) Ls[(ty + 16 * i) * LDD + tx + 16 * b] = a[i][b];
  __syncthreads();
  float* inv_out = invdiag + static_cast<long long>(k) * TS * TS;
  for (int idx = tid; idx < TS * TS; idx += 256) inv_out[idx] = Ls[(idx >> 7) * LDD + (idx & 127)];
}

// --------------------------------------------------------------------------------------------
// fp32 -> fp16 hi/lo operands of the tensor-core updates
// --------------------------------------------------------------------------------------------
constexpr int HDR_FLOATS = 64;  // workspace header: [0] max bits, [1] operand scale 2^s, [2] 2^-2s

__global__ void absmax_kernel(const float* __restrict__ A, long long M, long long lda, int diag_only,
                              unsigned* __restrict__ maxbits) {
  float m = 0.f;
  if (diag_only == 1) {
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < M;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
      m = fmaxf(m, fabsf(A[i * lda + i]));
  } else if (diag_only == 2) {  // all `lda` columns of the M rows (a dense M x lda array)
    for (long long r = blockIdx.x; r < M; r += gridDim.x)
      for (long long c = threadIdx.x; c < lda; c += blockDim.x) m = fmaxf(m, fabsf(A[r * lda + c]));
  } else {  // lower triangle incl. diagonal, one row per block iteration
    for (long long r = blockIdx.x; r < M; r += gridDim.x)
      for (long long c = threadIdx.x; c <= r; c += blockDim.x) m = fmaxf(m, fabsf(A[r * lda + c]));
  }
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(maxbits, __float_as_uint(m));
}

// scale = power of two that brings the largest operand entry into [2^12, 2^13): hi = fp16(x scale) is
// far from overflow and lo = fp16(x scale - hi) stays a NORMAL fp16 for every |x| > 2^-26 max.
// take_sqrt: the entries are bounded by sqrt(max) (Cholesky factor of a matrix with that diagonal).
__global__ void make_scale_kernel(float* hdr, int take_sqrt) {
  float m = __uint_as_float(reinterpret_cast<unsigned*>(hdr)[0]);
  if (take_sqrt) m = sqrtf(m);
  int e = 0;
  if (m > 0.f && isfinite(m)) frexpf(m, &e);  // m = f 2^e, f in [0.5, 1)
  const float sc = ldexpf(1.f, 13 - e);
  hdr[1] = sc;
  hdr[2] = 1.f / (sc * sc);
}

// fp32 operand -> K-concatenated fp16 pairs, W = K extent of the fp32 operand (multiple of 64):
//   outA[r] = [ lo | hi | hi ],  outB[r] = [ hi | lo | hi ]   (row stride 3 W), x 2^s = hi + lo
// so that outA outB^T = lo hi + hi lo + hi hi.  The small terms come first: the truncating tensor-core
// accumulator then only sees W/16 steps at full magnitude.  One thread per 8 consecutive K elements.
// transpose = 0: operand row r, K index k  <-  src[r][k]   (the Cholesky panel).
// transpose = 1: operand row r, K index k  <-  src[k][r] masked to r <= k_g0 + k   (block row of L
//                used as L^T; k_g0 = global row of src[0]).
__global__ void split_f16_kernel(const float* __restrict__ src, long long ld, long long rows, int W, int transpose,
                                 long long k_g0, const float* __restrict__ hdr, __half* __restrict__ outA,
                                 __half* __restrict__ outB) {
  const float sc = hdr[1];
  const int kgroups = W / 8;
  const long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  long long r;
  int kg;
  if (transpose) {  // consecutive threads -> consecutive operand rows (= consecutive src columns)
    kg = static_cast<int>(gid / rows);
    r = gid % rows;
    if (kg >= kgroups) return;
  } else {
    r = gid / kgroups;
    kg = static_cast<int>(gid % kgroups);
    if (r >= rows) return;
  }
  float x[8];
  if (transpose) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const long long k = kg * 8 + i;
      x[i] = (r <= k_g0 + k) ? src[k * ld + r] : 0.f;  // L[k_g0 + k][r], lower triangle only
    }
  } else {
    const float4 v0 = *reinterpret_cast<const float4*>(src + r * ld + kg * 8);
    const float4 v1 = *reinterpret_cast<const float4*>(src + r * ld + kg * 8 + 4);
    x[0] = v0.x; x[1] = v0.y; x[2] = v0.z; x[3] = v0.w;
    x[4] = v1.x; x[5] = v1.y; x[6] = v1.z; x[7] = v1.w;
  }
  __align__(16) __half hi[8];
  __align__(16) __half lo[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float v = x[i] * sc;
    hi[i] = __float2half_rn(v);
    lo[i] = __float2half_rn(v - __half2float(hi[i]));
  }
  const uint4 H = *reinterpret_cast<const uint4*>(hi), L = *reinterpret_cast<const uint4*>(lo);
  __half* a = outA + r * (3ll * W) + kg * 8;
  __half* b = outB + r * (3ll * W) + kg * 8;
  *reinterpret_cast<uint4*>(a) = L;
  *reinterpret_cast<uint4*>(a + W) = H;
  *reinterpret_cast<uint4*>(a + 2 * W) = H;
  *reinterpret_cast<uint4*>(b) = H;
  *reinterpret_cast<uint4*>(b + W) = L;
  *reinterpret_cast<uint4*>(b + 2 * W) = H;
}

// The diagonal of a tensor-core update is a sum of squares, where the truncating accumulator is
// biased; it decides positive definiteness, so it is computed here in fp32 and written back over the
// tensor-core result afterwards.
//   d[i] = C[i][i] + coef * sum_k P[i][k]^2      P = panel rows, W columns; one warp per row
__global__ void exact_diag_kernel(const float* __restrict__ C, long long ldc, const float* __restrict__ src,
                                  long long ld, long long rows, int W, float coef, float* __restrict__ dsave) {
  const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= rows) return;
  float s = 0.f;
  for (int k = lane; k < W; k += 32) {
    const float v = src[i * ld + k];
    s = fmaf(v, v, s);
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) dsave[i] = fmaf(coef, s, C[i * ldc + i]);
}
//   d[i] = gamma * G[i][i] + coef * sum_k L[k][i]^2 over the W slab rows with k_g0 + k >= i; one
//   thread per column i (consecutive threads read consecutive floats of a row)
__global__ void exact_diag_t_kernel(const float* __restrict__ G, long long ldg, const float* __restrict__ src,
                                    long long ld, long long cols, int W, long long k_g0, float gamma, float coef,
                                    float* __restrict__ dsave) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= cols) return;
  float s0 = 0.f, s1 = 0.f;
  long long kfirst = i - k_g0;
  if (kfirst < 0) kfirst = 0;
  int k = static_cast<int>(kfirst);
  for (; k + 1 < W; k += 2) {
    const float v0 = src[static_cast<long long>(k) * ld + i];
    const float v1 = src[static_cast<long long>(k + 1) * ld + i];
    s0 = fmaf(v0, v0, s0);
    s1 = fmaf(v1, v1, s1);
  }
  if (k < W) {
    const float v0 = src[static_cast<long long>(k) * ld + i];
    s0 = fmaf(v0, v0, s0);
  }
  const float old = gamma != 0.f ? gamma * G[i * ldg + i] : 0.f;
  dsave[i] = fmaf(coef, s0 + s1, old);
}

__global__ void restore_diag_kernel(float* __restrict__ C, long long ldc, long long rows,
                                    const float* __restrict__ dsave) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < rows) C[i * ldc + i] = dsave[i];
}

// the tensor-core accumulator truncates: an all-positive sum over W/16 steps at growing magnitude comes
// out low by ~0.8e-7 per step (calibrated in kmv.cu); the update is rescaled by the matching gain
static inline float tc_gain(int W) { return 1.f + 0.8e-7f * static_cast<float>(W / 16); }

// --------------------------------------------------------------------------------------------
// small elementwise / reduction kernels
// --------------------------------------------------------------------------------------------
__global__ void add_diag_kernel(float* A, long long M, long long lda, float v) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < M) A[i * lda + i] += v;
}

}  // namespace la
}  // namespace sab

using namespace sab;
using namespace sab::la;

#define SAB_REQUIRE_MAT(M, ld, ptr)                                                                    \
  SAB_REQUIRE((M) > 0 && (M) % TS == 0 && (ld) >= (M) && (ld) % 4 == 0 &&                              \
                  (reinterpret_cast<uintptr_t>(ptr) & 15) == 0,                                       \
              "matrix order must be a positive multiple of 128 with ld %% 4 == 0 and 16-byte aligned " \
              "storage (M=%lld ld=%lld)",                                                             \
              (long long)(M), (long long)(ld))

static bool use_tensor_path() {
  return option_int("SAB_LINALG_SIMT", 0) == 0;
}

// width (in 128-blocks) of the outer panels of the two-level factorisation: the trailing matrix is
// read and written once per outer panel, so wider panels move the update from HBM-bound to tensor-bound
static int outer_blocks() {
  int nb = option_int("SAB_CHOL_NB", 512);
  if (nb < TS) nb = TS;
  if (nb > 1024) nb = 1024;
  return nb / TS;
}

extern "C" int64_t sa_linalg_workspace_bytes(int64_t M) {
  if (M <= 0) return -1;
  const int64_t W = static_cast<int64_t>(outer_blocks()) * TS;
  return HDR_FLOATS * 4 + M * 4 + 2 * M * 3 * W * 2 + 256;
}

#define SAB_REQUIRE_WS(M, ws, bytes)                                                                  \
  SAB_REQUIRE((ws) != nullptr && (reinterpret_cast<uintptr_t>(ws) & 255) == 0 &&                      \
                  (bytes) >= sa_linalg_workspace_bytes(M),                                            \
              "workspace must be 256-byte aligned and hold sa_linalg_workspace_bytes(M) = %lld bytes", \
              (long long)sa_linalg_workspace_bytes(M))

struct LinalgWs {
  float* hdr;
  float* dsave;
  __half* opA;
  __half* opB;
};
static LinalgWs carve(void* workspace, int64_t M) {
  LinalgWs w;
  w.hdr = static_cast<float*>(workspace);
  w.dsave = w.hdr + HDR_FLOATS;
  uintptr_t p = reinterpret_cast<uintptr_t>(w.dsave + M);
  p = (p + 255) & ~static_cast<uintptr_t>(255);
  w.opA = reinterpret_cast<__half*>(p);
  w.opB = w.opA + M * 3 * static_cast<int64_t>(outer_blocks()) * TS;
  return w;
}

// ---- building blocks of the two-level Cholesky.  sa_potrf strings them together on one GPU; the
// row-sharded solver (sobolev_alignment_b200/distributed.py) deals the outer panels out over the ranks:
// the owner factors a panel (sa_potrf_panel), the panel travels by NCCL broadcast, every rank applies it to
// the outer panels it owns (sa_potrf_split once, sa_potrf_update per owned panel).
static int potrf_begin(cudaStream_t st, const float* A, int64_t M, int64_t lda, int* info, const LinalgWs& w,
                       bool tensor) {
  SAB_CHECK_CUDA(cudaMemsetAsync(info, 0, sizeof(int), st));
  if (tensor) {
    // |L_ij| <= sqrt(max_i A_ii): one power-of-two operand scale for the whole factorisation
    SAB_CHECK_CUDA(cudaMemsetAsync(w.hdr, 0, HDR_FLOATS * 4, st));
    absmax_kernel<<<static_cast<int>(ceil_div(M, 256)), 256, 0, st>>>(A, M, lda, 1,
                                                                      reinterpret_cast<unsigned*>(w.hdr));
    make_scale_kernel<<<1, 1, 0, st>>>(w.hdr, 1);
  }
  SAB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// factor the outer panel: block columns [k0, kend) x all rows below, 128-wide fp32 SIMT steps.
// inner_limit = last block column (exclusive) the in-panel updates reach: kend on the tensor path (the
// rest is the trailing update), nblk on the SIMT path (classic right-looking).
static int potrf_panel(cudaStream_t st, float* A, int64_t lda, float* invdiag, int* info, int nblk, int k0,
                       int kend, int inner_limit) {
  const int dsmem = (TS * LDD + 4 * TS) * static_cast<int>(sizeof(float));
  SAB_CHECK_CUDA(cudaFuncSetAttribute(diag_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dsmem));
  for (int k = k0; k < kend; ++k) {
    diag_factor_kernel<<<1, 256, dsmem, st>>>(A, lda, k, invdiag, info);
    const int rem = nblk - k - 1;
    if (rem == 0) break;
    sgemm_kernel<MODE_PANEL><<<rem, 256, 0, st>>>(A, lda, invdiag + static_cast<long long>(k) * TS * TS, TS, k,
                                                  nblk, 0.f, 0.f);
    const int jlim = inner_limit - k - 1;  // block columns right of k still inside the panel
    if (jlim >= rem)
      sgemm_kernel<MODE_SYRK><<<rem * (rem + 1) / 2, 256, 0, st>>>(A, lda, nullptr, 0, k, nblk,
                                                                   static_cast<float>(jlim), 0.f);
    else if (jlim > 0)
      sgemm_kernel<MODE_SYRK><<<rem * jlim, 256, 0, st>>>(A, lda, nullptr, 0, k, nblk, static_cast<float>(jlim), 0.f);
  }
  SAB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// fp16 hi/lo operands of a finished panel: `panel` = its rows BELOW the panel's diagonal blocks
// ([rows x W] fp32, row stride ldp), for the tensor-core trailing update
static int potrf_split(cudaStream_t st, const float* panel, int64_t ldp, long long rows, int W, const LinalgWs& w) {
  split_f16_kernel<<<static_cast<int>(ceil_div(rows * (W / 8), 256)), 256, 0, st>>>(panel, ldp, rows, W, 0, 0, w.hdr,
                                                                                 w.opA, w.opB);
  SAB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// C -= P P^T restricted to the block columns [c0, c1) (and the rows from c0 down) of the trailing matrix that
// starts at block kend; P = the split panel (row 0 of `panel` / the operands = global block row kend).
static int potrf_update(void* stream, float* A, int64_t lda, int nblk, int kend, int c0, int c1, const float* panel,
                        int64_t ldp, int W, const LinalgWs& w) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long off = static_cast<long long>(c0 - kend) * TS;      // first operand row of this update
  const long long ncols = static_cast<long long>(c1 - c0) * TS;
  const long long nrows = static_cast<long long>(nblk - c0) * TS;
  float* C = A + static_cast<long long>(c0) * TS * (lda + 1);
  // the diagonal of the update is a sum of squares: exact in fp32 (the tensor-core accumulator is biased there)
  exact_diag_kernel<<<static_cast<int>(ceil_div(ncols * 32, 256)), 256, 0, st>>>(C, lda, panel + off * ldp, ldp, ncols,
                                                                             W, -1.f, w.dsave + off);
  SAB_CHECK_CUDA(cudaGetLastError());
  if (int rc = tc_gemm_nt(stream, w.opA + off * 3 * W, nrows, 3 * W, w.opB + off * 3 * W, ncols, 3 * W, 3 * W, C, lda,
                          -tc_gain(W), w.hdr + 2, 1.f, 1))
    return rc;
  restore_diag_kernel<<<static_cast<int>(ceil_div(ncols, 256)), 256, 0, st>>>(C, lda, ncols, w.dsave + off);
  SAB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int sa_potrf(void* stream, float* A, int64_t M, int64_t lda, float* invdiag, int* info,
                        void* workspace, int64_t workspace_bytes) {
  SAB_REQUIRE_MAT(M, lda, A);
  SAB_REQUIRE(invdiag != nullptr && info != nullptr, "invdiag / info must not be NULL");
  SAB_REQUIRE_WS(M, workspace, workspace_bytes);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nblk = static_cast<int>(M / TS);
  const bool tensor = use_tensor_path();
  const int nbb = tensor ? outer_blocks() : nblk;  // SIMT: classic right-looking (one "outer panel")
  const LinalgWs w = carve(workspace, M);
  if (int rc = potrf_begin(st, A, M, lda, info, w, tensor)) return rc;
  for (int k0 = 0; k0 < nblk; k0 += nbb) {
    const int kend = k0 + nbb < nblk ? k0 + nbb : nblk;
    if (int rc = potrf_panel(st, A, lda, invdiag, info, nblk, k0, kend, tensor ? kend : nblk)) return rc;
    // ---- trailing update C -= P P^T on the tensor cores, P = the finished panel below row kend*128
    if (tensor && kend < nblk) {
      const long long rows = static_cast<long long>(nblk - kend) * TS;
      const int W = (kend - k0) * TS;
      const float* panel = A + static_cast<long long>(kend) * TS * lda + static_cast<long long>(k0) * TS;
      if (int rc = potrf_split(st, panel, lda, rows, W, w)) return rc;
      if (int rc = potrf_update(stream, A, lda, nblk, kend, kend, nblk, panel, lda, W, w)) return rc;
    }
  }
  return 0;
}

extern "C" int sa_potrf_outer_blocks(void) { return use_tensor_path() ? outer_blocks() : 0; }

extern "C" int sa_potrf_begin(void* stream, const float* A, int64_t M, int64_t lda, int* info, void* workspace,
                              int64_t workspace_bytes) {
  SAB_REQUIRE_MAT(M, lda, A);
  SAB_REQUIRE(info != nullptr, "info must not be NULL");
  SAB_REQUIRE(use_tensor_path(), "the panel-wise factorisation needs the tensor-core path");
  SAB_REQUIRE_WS(M, workspace, workspace_bytes);
  return potrf_begin(static_cast<cudaStream_t>(stream), A, M, lda, info, carve(workspace, M), true);
}

extern "C" int sa_potrf_panel(void* stream, float* A, int64_t M, int64_t lda, float* invdiag, int* info, int k0,
                              int kend) {
  SAB_REQUIRE_MAT(M, lda, A);
  const int nblk = static_cast<int>(M / TS);
  SAB_REQUIRE(invdiag != nullptr && info != nullptr && 0 <= k0 && k0 < kend && kend <= nblk &&
                  kend - k0 <= 1024 / TS,
              "bad panel [%d, %d) of %d blocks", k0, kend, nblk);
  return potrf_panel(static_cast<cudaStream_t>(stream), A, lda, invdiag, info, nblk, k0, kend, kend);
}

extern "C" int sa_potrf_split(void* stream, int64_t M, const float* panel, int64_t ldp, int64_t rows, int W,
                              void* workspace, int64_t workspace_bytes) {
  SAB_REQUIRE(M > 0 && M % TS == 0 && panel != nullptr && rows > 0 && rows % TS == 0 && rows < M && W > 0 &&
                  W % TS == 0 && W <= outer_blocks() * TS && ldp >= W && ldp % 4 == 0 &&
                  (reinterpret_cast<uintptr_t>(panel) & 15) == 0,
              "bad panel operand (rows=%lld W=%d ldp=%lld)", (long long)rows, W, (long long)ldp);
  SAB_REQUIRE_WS(M, workspace, workspace_bytes);
  return potrf_split(static_cast<cudaStream_t>(stream), panel, ldp, rows, W, carve(workspace, M));
}

extern "C" int sa_potrf_update(void* stream, float* A, int64_t M, int64_t lda, int kend, int c0, int c1,
                               const float* panel, int64_t ldp, int W, void* workspace, int64_t workspace_bytes) {
  SAB_REQUIRE_MAT(M, lda, A);
  const int nblk = static_cast<int>(M / TS);
  SAB_REQUIRE(panel != nullptr && 0 < kend && kend <= c0 && c0 < c1 && c1 <= nblk && W > 0 && W % TS == 0 &&
                  W <= outer_blocks() * TS && ldp >= W,
              "bad update: kend=%d columns [%d, %d) of %d blocks, W=%d", kend, c0, c1, nblk, W);
  SAB_REQUIRE_WS(M, workspace, workspace_bytes);
  return potrf_update(stream, A, lda, nblk, kend, c0, c1, panel, ldp, W, carve(workspace, M));
}

// All the outer panels one rank owns in ONE call (the per-panel Python round trip -- ~40 us of ctypes per update --
// was what bounded the distributed factorisation): panels first, first + stride, ... (outer-panel indices, each
// `nb` blocks wide) except `skip` (the look-ahead panel, already brought up to date).
extern "C" int sa_potrf_update_cyclic(void* stream, float* A, int64_t M, int64_t lda, int kend, int nb, int first,
                                      int stride, int skip, const float* panel, int64_t ldp, int W, void* workspace,
                                      int64_t workspace_bytes) {
  SAB_REQUIRE_MAT(M, lda, A);
  const int nblk = static_cast<int>(M / TS);
  SAB_REQUIRE(panel != nullptr && 0 < kend && kend <= nblk && nb > 0 && nb <= outer_blocks() && first >= 0 && stride > 0 &&
                  W > 0 && W % TS == 0 && W <= outer_blocks() * TS && ldp >= W,
              "bad cyclic update: kend=%d nb=%d first=%d stride=%d W=%d", kend, nb, first, stride, W);
  SAB_REQUIRE_WS(M, workspace, workspace_bytes);
  const LinalgWs w = carve(workspace, M);
  for (int j = first; j * nb < nblk; j += stride) {
    if (j == skip) continue;
    const int c0 = j * nb, c1 = c0 + nb < nblk ? c0 + nb : nblk;
    SAB_REQUIRE(c0 >= kend, "panel %d starts before the trailing matrix (block %d)", j, kend);
    if (int rc = potrf_update(stream, A, lda, nblk, kend, c0, c1, panel, ldp, W, w)) return rc;
  }
  return 0;
}

extern "C" int sa_ltl_part(void* stream, const float* L, int64_t M, int64_t ldl, float* G, int64_t ldg,
                           float alpha, float beta, void* workspace, int64_t workspace_bytes, int part,
                           int nparts) {
  SAB_REQUIRE(nparts >= 1 && part >= 0 && part < nparts, "bad part %d of %d", part, nparts);
  SAB_REQUIRE_MAT(M, ldl, L);
  SAB_REQUIRE_MAT(M, ldg, G);
  SAB_REQUIRE(L != G, "sa_ltl is out of place");
  SAB_REQUIRE_WS(M, workspace, workspace_bytes);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nblk = static_cast<int>(M / TS);
  SAB_REQUIRE(nblk <= 2048, "matrix order %lld too large", (long long)M);
  if (!use_tensor_path()) {
    SAB_REQUIRE(nparts == 1, "the fp32 SIMT path computes L^T L in one piece");
    sgemm_kernel<MODE_LTL><<<nblk * (nblk + 1) / 2, 256, 0, st>>>(G, ldg, L, ldl, 0, nblk, alpha, beta);
    SAB_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  const LinalgWs w = carve(workspace, M);
  const int nbb = outer_blocks();
  SAB_CHECK_CUDA(cudaMemsetAsync(w.hdr, 0, HDR_FLOATS * 4, st));
  absmax_kernel<<<1184, 256, 0, st>>>(L, M, ldl, 0, reinterpret_cast<unsigned*>(w.hdr));
  make_scale_kernel<<<1, 1, 0, st>>>(w.hdr, 0);
  // G = alpha sum_b L_b^T L_b over slabs of block rows; slab b = rows [r0, r1) touches G[0 .. r1)^2.
  // Starting with the last slab, the first update of every tile is the one with gamma = 0.
  // nparts > 1: only the slabs b = part (mod nparts) are accumulated INTO G (which the caller has zeroed) and
  // no diagonal term is added: the parts of the ranks are summed by an all-reduce, then beta goes on top.
  const int nslab = static_cast<int>(ceil_div(nblk, nbb));
  for (int b = nslab - 1; b >= 0; --b) {
    if (nparts > 1 && (nslab - 1 - b) % nparts != part) continue;  // counted from the last (largest) slab
    const int b0 = b * nbb, b1 = (b + 1) * nbb < nblk ? (b + 1) * nbb : nblk;
    const long long rows = static_cast<long long>(b1) * TS;  // non-zero columns of the slab
    const int W = (b1 - b0) * TS;
    const float* slab = L + static_cast<long long>(b0) * TS * ldl;
    const float gamma = (nparts == 1 && b == nslab - 1) ? 0.f : 1.f;
    exact_diag_t_kernel<<<static_cast<int>(ceil_div(rows, 256)), 256, 0, st>>>(
        G, ldg, slab, ldl, rows, W, static_cast<long long>(b0) * TS, gamma, alpha, w.dsave);
    split_f16_kernel<<<static_cast<int>(ceil_div(rows * (W / 8), 256)), 256, 0, st>>>(
        slab, ldl, rows, W, 1, static_cast<long long>(b0) * TS, w.hdr, w.opA, w.opB);
    SAB_CHECK_CUDA(cudaGetLastError());
    if (int rc = tc_gemm_nt(stream, w.opA, rows, 3 * W, w.opB, rows, 3 * W, 3 * W, G, ldg, alpha * tc_gain(W),
                            w.hdr + 2, gamma, 1))
      return rc;
    restore_diag_kernel<<<static_cast<int>(ceil_div(rows, 256)), 256, 0, st>>>(G, ldg, rows, w.dsave);
  }
  if (nparts == 1) add_diag_kernel<<<static_cast<int>(ceil_div(M, 256)), 256, 0, st>>>(G, M, ldg, beta);
  SAB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int sa_ltl(void* stream, const float* L, int64_t M, int64_t ldl, float* G, int64_t ldg,
                      float alpha, float beta, void* workspace, int64_t workspace_bytes) {
  return sa_ltl_part(stream, L, M, ldl, G, ldg, alpha, beta, workspace, workspace_bytes, 0, 1);
}

extern "C" int sa_add_diag(void* stream, float* A, int64_t M, int64_t lda, float value) {
  SAB_REQUIRE(M >= 0 && lda >= M, "bad shape");
  if (M == 0) return 0;
  add_diag_kernel<<<static_cast<int>(ceil_div(M, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      A, M, lda, value);
  SAB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
