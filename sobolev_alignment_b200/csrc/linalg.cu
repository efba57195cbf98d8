// fp32 dense linear algebra of the Nystrom preconditioner and the CG vector updates.
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
//   sa_trsm    block forward / backward substitution; one launch per 128-row block step, each CTA
//              streams one 128x128 tile of L (coalesced 128-bit loads) against the M x dp rhs
//   sa_coldot / sa_cg_update / sa_cg_direction / sa_axpby   CG vector kernels, coalesced over the
//              row-major M x dp layout with warp-shuffle + shared-memory column reductions
//
// Replaces falkon `FalkonPreconditioner` (potrf, lauum, mul_triang, trsm) and
// `ConjugateGradient.solve` [external dependency], reached from
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
// block triangular solves
// --------------------------------------------------------------------------------------------
// The 128x128 tile product out = op(T) x runs on the legacy tensor path (mma.sync m16n8k8 TF32) with
// the 3xTF32 split (a = a_hi + a_lo, b = b_hi + b_lo; a_hi b_hi + a_lo b_hi + a_hi b_lo), which keeps
// ~21 bits: the solve stays HBM / latency bound instead of shared-memory-broadcast bound.
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
  const float r = x - __uint_as_float(hi);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// shared-memory row strides chosen so that every fragment load is bank-conflict free
__host__ __device__ constexpr int trsm_ldt(bool trans) { return trans ? TS + 8 : TS + 4; }
__host__ __device__ constexpr int trsm_ldx(int nb8) { return nb8 == 1 ? 8 : (nb8 <= 3 ? 24 : 40); }

// warp w owns output rows [16w, 16w+16); acc[nb] is the m16n8 fragment of column block nb
template <int NB8, bool TRANS>
__device__ __forceinline__ void tile_times_rhs(const float* Ts, const float* xs, float (&acc)[NB8][4]) {
  constexpr int LT = trsm_ldt(TRANS);
  constexpr int LX = trsm_ldx(NB8);
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int m0 = (threadIdx.x >> 5) * 16;
  // three independent accumulator chains per column block (hi*hi, lo*hi, hi*lo) keep the dependent
  // mma.sync chain at 16 instead of 48 on the critical path of the diagonal CTA
  float c1[NB8][4], c2[NB8][4];
#pragma unroll
  for (int nb = 0; nb < NB8; ++nb)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[nb][i] = c1[nb][i] = c2[nb][i] = 0.f;
#pragma unroll 4
  for (int k0 = 0; k0 < TS; k0 += 8) {
    float a[4];
    if (TRANS) {
      a[0] = Ts[(k0 + t) * LT + m0 + g];
      a[1] = Ts[(k0 + t) * LT + m0 + g + 8];
      a[2] = Ts[(k0 + t + 4) * LT + m0 + g];
      a[3] = Ts[(k0 + t + 4) * LT + m0 + g + 8];
    } else {
      a[0] = Ts[(m0 + g) * LT + k0 + t];
      a[1] = Ts[(m0 + g + 8) * LT + k0 + t];
      a[2] = Ts[(m0 + g) * LT + k0 + t + 4];
      a[3] = Ts[(m0 + g + 8) * LT + k0 + t + 4];
    }
    uint32_t ahi[4], alo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split_tf32(a[i], ahi[i], alo[i]);
#pragma unroll
    for (int nb = 0; nb < NB8; ++nb) {
      uint32_t bhi[2], blo[2];
      split_tf32(xs[(k0 + t) * LX + nb * 8 + g], bhi[0], blo[0]);
      split_tf32(xs[(k0 + t + 4) * LX + nb * 8 + g], bhi[1], blo[1]);
      mma_tf32(acc[nb], ahi, bhi);
      mma_tf32(c1[nb], alo, bhi);
      mma_tf32(c2[nb], ahi, blo);
    }
  }
#pragma unroll
  for (int nb = 0; nb < NB8; ++nb)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[nb][i] += c1[nb][i] + c2[nb][i];
}

// ---- fp16 hi/lo tile product (half the MMAs of the 3xTF32 form; the chained solve is bound by their throughput)
// x is converted ONCE per block into packed fp16 pairs, transposed to [column][k] so that a B fragment is one
// 32-bit shared load; the tile is split in registers.  Scales are powers of two: the tile by `tscale` (from
// the largest entry of the whole factor / inverse blocks), x by the block's own largest entry.
constexpr int XH_LD = TS + 8;  // halves per column of the converted x (bank-conflict free fragment loads)

__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void split_h2(float v0, float v1, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(v0, v1);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// Converts the x block in xs [128][LX] (fp32) into xh / xl [NB8*8][XH_LD] (fp16 hi / lo of x * 2^e) and returns
// 2^-e.  Block-wide: contains two __syncthreads; `red` is 8 floats of shared scratch.
template <int NB8>
__device__ __forceinline__ float convert_x(const float* xs, __half* xh, __half* xl, float* red) {
  constexpr int LX = trsm_ldx(NB8);
  constexpr int NC = NB8 * 8;
  float m = 0.f;
  for (int idx = threadIdx.x; idx < TS * NC; idx += 256) m = fmaxf(m, fabsf(xs[(idx / NC) * LX + idx % NC]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  m = red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]);
  int e = 0;
  if (m > 0.f && isfinite(m)) frexpf(m, &e);       // m = f 2^e, f in [0.5, 1)
  const float sc = ldexpf(1.f, 12 - e), inv = ldexpf(1.f, e - 12);
  for (int idx = threadIdx.x; idx < TS * NC; idx += 256) {
    const int k = idx / NC, c = idx % NC;
    const float v = xs[k * LX + c] * sc;
    const __half h = __float2half_rn(v);
    xh[c * XH_LD + k] = h;
    xl[c * XH_LD + k] = __float2half_rn(v - __half2float(h));
  }
  __syncthreads();
  return inv;
}

// acc = (Ts x) with Ts scaled by tscale (power of two) and x given as the converted hi / lo halves; the caller
// multiplies by 1 / (tscale * 2^e).  Same fragment ownership as tile_times_rhs.
template <int NB8, bool TRANS>
__device__ __forceinline__ void tile_times_xh(const float* Ts, const __half* xh, const __half* xl, float tscale,
                                              float (&acc)[NB8][4]) {
  constexpr int LT = trsm_ldt(TRANS);
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int m0 = (threadIdx.x >> 5) * 16;
  float c1[NB8][4];
#pragma unroll
  for (int nb = 0; nb < NB8; ++nb)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[nb][i] = c1[nb][i] = 0.f;
#pragma unroll 2
  for (int k0 = 0; k0 < TS; k0 += 16) {
    // A fragment (m16 x k16): a0 (row g, k 2t..2t+1), a1 (row g+8, same k), a2 / a3: k + 8
    float v[4][2];
    if (TRANS) {
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int k = k0 + 2 * t + 8 * h + q;
          v[2 * h][q] = Ts[k * LT + m0 + g];
          v[2 * h + 1][q] = Ts[k * LT + m0 + g + 8];
        }
    } else {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float2 r0 = *reinterpret_cast<const float2*>(Ts + (m0 + g) * LT + k0 + 2 * t + 8 * h);
        const float2 r1 = *reinterpret_cast<const float2*>(Ts + (m0 + g + 8) * LT + k0 + 2 * t + 8 * h);
        v[2 * h][0] = r0.x; v[2 * h][1] = r0.y;
        v[2 * h + 1][0] = r1.x; v[2 * h + 1][1] = r1.y;
      }
    }
    uint32_t ahi[4], alo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split_h2(v[i][0] * tscale, v[i][1] * tscale, ahi[i], alo[i]);
#pragma unroll
    for (int nb = 0; nb < NB8; ++nb) {
      const int col = nb * 8 + g;
      uint32_t bhi[2], blo[2];
      bhi[0] = *reinterpret_cast<const uint32_t*>(xh + col * XH_LD + k0 + 2 * t);
      bhi[1] = *reinterpret_cast<const uint32_t*>(xh + col * XH_LD + k0 + 2 * t + 8);
      blo[0] = *reinterpret_cast<const uint32_t*>(xl + col * XH_LD + k0 + 2 * t);
      blo[1] = *reinterpret_cast<const uint32_t*>(xl + col * XH_LD + k0 + 2 * t + 8);
      mma_f16(acc[nb], ahi, bhi);
      mma_f16(c1[nb], alo, bhi);
      mma_f16(c1[nb], ahi, blo);
    }
  }
#pragma unroll
  for (int nb = 0; nb < NB8; ++nb)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[nb][i] += c1[nb][i];
}

__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(
                   static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst))),
               "l"(gmem_src)
               : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// all 16 128-bit requests of a thread are in flight at once (the solve is latency bound)
template <bool TRANS>
__device__ __forceinline__ void load_tile_128_async(const float* __restrict__ src, long long ld, float* Ts) {
  constexpr int LT = trsm_ldt(TRANS);
#pragma unroll
  for (int i = 0; i < TS * TS / 4 / 256; ++i) {
    const int idx = threadIdx.x + i * 256;
    const int r = idx >> 5, cq = idx & 31;
    cp_async_16(Ts + r * LT + cq * 4, src + static_cast<long long>(r) * ld + cq * 4);
  }
}

// Forward step k (TRANS = false): CTA i = k + blockIdx.x:  b_i -= L(i, k-1) x_{k-1};  CTA i == k then
// solves x_k = inv(L_kk) b_k.  Backward step k (TRANS = true): CTA j = blockIdx.x <= k:
// b_j -= L(k+1, j)^T x_{k+1};  CTA j == k then solves x_k = inv(L_kk)^T b_k.
template <int DP, bool TRANS>
__global__ void __launch_bounds__(256)
trsm_step_kernel(const float* __restrict__ L, long long ldl, const float* __restrict__ invdiag,
                 float* __restrict__ Bm, int k, int nblk) {
  constexpr int NB8 = (DP + 7) / 8;
  constexpr int LT = trsm_ldt(TRANS);
  constexpr int LX = trsm_ldx(NB8);
  extern __shared__ __align__(16) float tsm[];
  float* Ts = tsm;            // [128][LT]
  float* xs = tsm + TS * LT;  // [128][LX]
  const int blk = TRANS ? static_cast<int>(blockIdx.x) : k + static_cast<int>(blockIdx.x);
  const int src = TRANS ? k + 1 : k - 1;  // block whose solution is being eliminated
  const bool has_update = TRANS ? (k + 1 < nblk) : (k > 0);
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int m0 = (threadIdx.x >> 5) * 16;
  // Programmatic dependent launch: let the next step start its prologue now; everything that does
  // not depend on the previous step (the L tile, the inverted diagonal block) is requested before
  // waiting for that step to complete.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  // padded columns of the rhs tile must be finite
  if (LX > DP)
    for (int idx = threadIdx.x; idx < TS * (LX - DP); idx += 256)
      xs[(idx / (LX - DP)) * LX + DP + idx % (LX - DP)] = 0.f;
  if (has_update) {
    const float* tile = TRANS ? L + static_cast<long long>(src) * TS * ldl + static_cast<long long>(blk) * TS
                              : L + static_cast<long long>(blk) * TS * ldl + static_cast<long long>(src) * TS;
    load_tile_128_async<TRANS>(tile, ldl, Ts);
  }
  float4 inv[TS * TS / 4 / 256];
  if (blk == k) {
    const float* isrc = invdiag + static_cast<long long>(k) * TS * TS;
#pragma unroll
    for (int i = 0; i < TS * TS / 4 / 256; ++i)
      inv[i] = __ldg(reinterpret_cast<const float4*>(isrc) + threadIdx.x + i * 256);
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");  // the previous step's x / b are now visible

  // this thread's fragment of the block of b: rows m0+g and m0+g+8, columns nb*8 + 2t, +1
  float* b0 = Bm + (static_cast<long long>(blk) * TS + m0 + g) * DP;
  float* b1 = b0 + 8 * DP;
  float bfrag[NB8][4];
#pragma unroll
  for (int nb = 0; nb < NB8; ++nb) {
    const int c = nb * 8 + 2 * t;
    float2 v0 = make_float2(0.f, 0.f), v1 = make_float2(0.f, 0.f);
    if (c < DP) {
      v0 = *reinterpret_cast<const float2*>(b0 + c);
      v1 = *reinterpret_cast<const float2*>(b1 + c);
    }
    bfrag[nb][0] = v0.x; bfrag[nb][1] = v0.y; bfrag[nb][2] = v1.x; bfrag[nb][3] = v1.y;
  }
  if (has_update) {
    const float* xsrc = Bm + static_cast<long long>(src) * TS * DP;
    for (int idx = threadIdx.x; idx < TS * (DP / 4); idx += 256)
      cp_async_16(xs + (idx / (DP / 4)) * LX + (idx % (DP / 4)) * 4, xsrc + idx * 4);
  }
  if (has_update) {
    cp_async_wait_all();
    __syncthreads();
    float acc[NB8][4];
    tile_times_rhs<NB8, TRANS>(Ts, xs, acc);
#pragma unroll
    for (int nb = 0; nb < NB8; ++nb)
#pragma unroll
      for (int i = 0; i < 4; ++i) bfrag[nb][i] -= acc[nb][i];
  }
  if (blk == k) {
    __syncthreads();  // everyone is done with Ts / xs
#pragma unroll
    for (int i = 0; i < TS * TS / 4 / 256; ++i) {
      const int idx = threadIdx.x + i * 256;
      *reinterpret_cast<float4*>(Ts + (idx >> 5) * LT + (idx & 31) * 4) = inv[i];
    }
#pragma unroll
    for (int nb = 0; nb < NB8; ++nb) {
      const int c = nb * 8 + 2 * t;
      xs[(m0 + g) * LX + c] = bfrag[nb][0];
      xs[(m0 + g) * LX + c + 1] = bfrag[nb][1];
      xs[(m0 + g + 8) * LX + c] = bfrag[nb][2];
      xs[(m0 + g + 8) * LX + c + 1] = bfrag[nb][3];
    }
    __syncthreads();
    float acc[NB8][4];
    tile_times_rhs<NB8, TRANS>(Ts, xs, acc);
#pragma unroll
    for (int nb = 0; nb < NB8; ++nb)
#pragma unroll
      for (int i = 0; i < 4; ++i) bfrag[nb][i] = acc[nb][i];
  }
#pragma unroll
  for (int nb = 0; nb < NB8; ++nb) {
    const int c = nb * 8 + 2 * t;
    if (c < DP) {
      *reinterpret_cast<float2*>(b0 + c) = make_float2(bfrag[nb][0], bfrag[nb][1]);
      *reinterpret_cast<float2*>(b1 + c) = make_float2(bfrag[nb][2], bfrag[nb][3]);
    }
  }
}

// --------------------------------------------------------------------------------------------
// Single-launch block triangular solve.  CTA c solves block row blk(c) (forward: c, backward:
// nblk-1-c): it streams the tiles L(blk, src) of the blocks solved before it through a two-slot
// cp.async ring, eliminates each x_src once it has been published (flag in global memory, release /
// acquire at gpu scope), finally multiplies by the pre-inverted diagonal block and publishes its own
// x.  When the NEXT dependency is already published -- always the case while a CTA works off the
// blocks solved before it became resident -- its x is prefetched together with its tile, so such
// steps cost one tile product and no exposed memory latency.  Logical CTA indices come from an
// atomic ticket, so a CTA only ever waits for CTAs that have already started (no dependence on the
// hardware's dispatch order).
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// One solve, or two pipelined ones: job 1 solves  op(L1) x1 = x0 + lambda V  where x0 is job 0's
// solution.  Both chains run in the same block order, so block c of job 1 starts as soon as job 0 has
// published block c: the pair costs one chain latency instead of two (the FALKON operator applies
// A^-1 then T^-1, and T^-T then A^-T, back to back).
struct TrsmJob {
  const float* L;
  long long ldl;
  const float* invdiag;
  const float* rhs;  // job 0: == x (in place); job 1: job 0's x
  float* x;          // solution, M x DP
  const float* V;    // optional additive term of the rhs (NULL = none)
  float lambda;
  const float* W;    // optional pre-multiplied sub-diagonal tiles from sa_trsm_premul (NULL = none)
  const float* scales;  // behind the W tiles: power-of-two operand scales {L, inverse blocks, W} of the fp16 product
};
struct TrsmJobs {
  TrsmJob job[2];
  int njobs;
  long long* trace;  // developer tool (SAB_TRSM_TRACE_PTR): per block 8 globaltimer stamps of the last chain step
};

template <int DP, bool TRANS>
__global__ void __launch_bounds__(256)
trsm_chain_kernel(const TrsmJobs jobs, int nblk, int* __restrict__ sync) {
  constexpr int NB8 = (DP + 7) / 8;
  constexpr int LT = trsm_ldt(TRANS);
  constexpr int LX = trsm_ldx(NB8);
  extern __shared__ __align__(16) float tsm[];
  float* ring = tsm;                  // [2][128][LT]
  float* xring = tsm + 2 * TS * LT;   // [2][128][LX]
  __shared__ int s_ticket, s_next;
  if (threadIdx.x == 0) s_ticket = atomicAdd(sync, 1);
  __syncthreads();
  const int jid = s_ticket % jobs.njobs;  // tickets alternate between the jobs: both chains advance together
  const int c = s_ticket / jobs.njobs;    // c blocks of this job are solved before this one
  const TrsmJob& J = jobs.job[jid];
  const float* __restrict__ L = J.L;
  const long long ldl = J.ldl;
  const float* __restrict__ invdiag = J.invdiag;
  float* __restrict__ Bm = J.x;
  const int blk = TRANS ? nblk - 1 - c : c;
  int* flags = sync + 64 + jid * nblk;
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int m0 = (threadIdx.x >> 5) * 16;

  auto src_of = [&](int d) { return TRANS ? nblk - 1 - d : d; };
  auto issue_tile = [&](int d) {  // d < c: tile of dependency d; d == c: the inverted diagonal block
    float* dst = ring + (d & 1) * TS * LT;
    if (d < c) {
      const int src = src_of(d);
      const float* tile = TRANS ? L + static_cast<long long>(src) * TS * ldl + static_cast<long long>(blk) * TS
                                : L + static_cast<long long>(blk) * TS * ldl + static_cast<long long>(src) * TS;
      load_tile_128_async<TRANS>(tile, ldl, dst);
    } else {
      load_tile_128_async<TRANS>(invdiag + static_cast<long long>(blk) * TS * TS, TS, dst);
    }
  };
  auto issue_x = [&](int d) {
    const float* xsrc = Bm + static_cast<long long>(src_of(d)) * TS * DP;
    float* xs = xring + (d & 1) * TS * LX;
    for (int idx = threadIdx.x; idx < TS * (DP / 4); idx += 256)
      cp_async_16(xs + (idx / (DP / 4)) * LX + (idx % (DP / 4)) * 4, xsrc + idx * 4);
  };

  // padded columns of the rhs tiles must be finite
  if (LX > DP)
    for (int idx = threadIdx.x; idx < 2 * TS * (LX - DP); idx += 256)
      xring[(idx / (LX - DP)) * LX + DP + idx % (LX - DP)] = 0.f;
  issue_tile(0);
  cp_async_commit();

  auto wait_flag = [&](const int* f, int what) {  // bounded spin: a bug must trap, not hang the GPU
    unsigned spins = 0;
    unsigned long long t0 = 0;
    while (ld_acquire_gpu(f) == 0) {
      if ((++spins & 0x3FFu) == 0) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        else if (now - t0 > 4000000000ull) {
          printf("sobolev_b200: trsm chain timed out waiting for block %d (job %d cta %d)\n", what, jid, c);
          __trap();
        }
      }
    }
  };
  // this thread's fragment of the block of b: rows m0+g and m0+g+8, columns nb*8 + 2t, +1
  const long long frow = static_cast<long long>(blk) * TS + m0 + g;
  float* b0 = Bm + frow * DP;
  float* b1 = b0 + 8 * DP;
  // bfrag accumulates  rhs - sum_k L(blk, k) x_k.  Job 0 reads its rhs right away; job 1's rhs is job 0's
  // solution of the same block, the LAST thing it needs: it is fetched after the elimination loop, so
  // the second chain trails the first by one step instead of serialising behind it.
  float bfrag[NB8][4];
  auto load_rhs = [&]() {
#pragma unroll
    for (int nb = 0; nb < NB8; ++nb) {
      const int cc = nb * 8 + 2 * t;
      float2 v0 = make_float2(0.f, 0.f), v1 = make_float2(0.f, 0.f);
      if (cc < DP) {
        v0 = __ldcg(reinterpret_cast<const float2*>(J.rhs + frow * DP + cc));
        v1 = __ldcg(reinterpret_cast<const float2*>(J.rhs + (frow + 8) * DP + cc));
        if (J.V != nullptr) {
          const float2 a0 = *reinterpret_cast<const float2*>(J.V + frow * DP + cc);
          const float2 a1 = *reinterpret_cast<const float2*>(J.V + (frow + 8) * DP + cc);
          v0.x = fmaf(J.lambda, a0.x, v0.x); v0.y = fmaf(J.lambda, a0.y, v0.y);
          v1.x = fmaf(J.lambda, a1.x, v1.x); v1.y = fmaf(J.lambda, a1.y, v1.y);
        }
      }
      bfrag[nb][0] += v0.x; bfrag[nb][1] += v0.y; bfrag[nb][2] += v1.x; bfrag[nb][3] += v1.y;
    }
  };
#pragma unroll
  for (int nb = 0; nb < NB8; ++nb)
#pragma unroll
    for (int i = 0; i < 4; ++i) bfrag[nb][i] = 0.f;
  if (jid == 0) load_rhs();

  auto stamp = [&](int slot) {
    if (jobs.trace != nullptr && jid == 0 && threadIdx.x == 0) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      jobs.trace[8 * blk + slot] = static_cast<long long>(now);
    }
  };
  bool have_x = false;  // x of the current dependency was prefetched by the previous iteration
  for (int d = 0; d < c; ++d) {
    if (threadIdx.x == 0) {
      if (d == c - 1) stamp(0);
      if (!have_x) wait_flag(flags + src_of(d), src_of(d));
      if (d == c - 1) stamp(1);
      s_next = (d + 1 < c) ? ld_acquire_gpu(flags + src_of(d + 1)) : 0;
    }
    __syncthreads();  // x_src is published; everyone is done with slot (d + 1) & 1 of both rings
    const bool nxt = s_next != 0;
    if (!have_x) {
      issue_x(d);
      cp_async_commit();
    }
    issue_tile(d + 1);  // next dependency's tile (or the inverted diagonal block) streams meanwhile ...
    if (nxt) issue_x(d + 1);  // ... with its x when that is already there
    cp_async_commit();
    cp_async_wait<1>();  // everything but that newest group: tile d and x_d have landed
    __syncthreads();
    if (d == c - 1) stamp(2);
    float acc[NB8][4];
    tile_times_rhs<NB8, TRANS>(ring + (d & 1) * TS * LT, xring + (d & 1) * TS * LX, acc);
    if (d == c - 1) stamp(3);
#pragma unroll
    for (int nb = 0; nb < NB8; ++nb)
#pragma unroll
      for (int i = 0; i < 4; ++i) bfrag[nb][i] -= acc[nb][i];
    have_x = nxt;
  }
  if (jid == 1) {
    if (threadIdx.x == 0) wait_flag(sync + 64 + blk, blk);  // job 0 has published this block
    __syncthreads();
    load_rhs();
  }
  cp_async_wait<0>();
  __syncthreads();  // inverted diagonal block landed; x ring free
  float* xs = xring;
#pragma unroll
  for (int nb = 0; nb < NB8; ++nb) {
    const int cc = nb * 8 + 2 * t;
    xs[(m0 + g) * LX + cc] = bfrag[nb][0];
    xs[(m0 + g) * LX + cc + 1] = bfrag[nb][1];
    xs[(m0 + g + 8) * LX + cc] = bfrag[nb][2];
    xs[(m0 + g + 8) * LX + cc + 1] = bfrag[nb][3];
  }
  __syncthreads();
  stamp(4);
  float acc[NB8][4];
  tile_times_rhs<NB8, TRANS>(ring + (c & 1) * TS * LT, xs, acc);
  stamp(5);
#pragma unroll
  for (int nb = 0; nb < NB8; ++nb) {
    const int cc = nb * 8 + 2 * t;
    if (cc < DP) {
      *reinterpret_cast<float2*>(b0 + cc) = make_float2(acc[nb][0], acc[nb][1]);
      *reinterpret_cast<float2*>(b1 + cc) = make_float2(acc[nb][2], acc[nb][3]);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    stamp(6);
    __threadfence();
    st_release_gpu(flags + blk, 1);
    stamp(7);
  }
}

// --------------------------------------------------------------------------------------------
// Pre-multiplied sub-diagonal tiles.  In the chained solve the block that has just been published is
// the LAST dependency of the next block, and two dependent tile products follow it (eliminate, then
// multiply by the inverted diagonal block): 5.4 of the 7.9 us per chain step.  With
//     Wf[c] = inv(L_cc) L(c, c-1)          (forward)        Wb[c] = L(c+1, c) inv(L_cc)   (backward, used transposed)
// the step becomes  x_c = inv(L_cc) (b_c - sum_{k < c-1} L(c,k) x_k)  -  Wf[c] x_{c-1}: everything but the last
// product is done before x_{c-1} arrives.
// --------------------------------------------------------------------------------------------
constexpr int PM_LDA = TS + 1;

__global__ void __launch_bounds__(256)
trsm_premul_kernel(const float* __restrict__ L, long long ldl, const float* __restrict__ invdiag, int nblk,
                   float* __restrict__ W) {
  extern __shared__ float pms[];
  float* As = pms;                 // [128][129]  left factor, row-major
  float* Bs = pms + TS * PM_LDA;   // [128][128]  right factor, row-major
  const int c = blockIdx.x, which = blockIdx.y;
  float* out = W + (static_cast<long long>(c) * 2 + which) * TS * TS;
  const bool valid = which == 0 ? c >= 1 : c + 1 < nblk;
  const int tid = threadIdx.x;
  if (!valid) {
    for (int idx = tid; idx < TS * TS; idx += 256) out[idx] = 0.f;
    return;
  }
  const float* inv = invdiag + static_cast<long long>(c) * TS * TS;
  const float* A = which == 0 ? inv : L + static_cast<long long>(c + 1) * TS * ldl + static_cast<long long>(c) * TS;
  const long long lda = which == 0 ? TS : ldl;
  const float* B = which == 0 ? L + static_cast<long long>(c) * TS * ldl + static_cast<long long>(c - 1) * TS : inv;
  const long long ldb = which == 0 ? ldl : TS;
  for (int idx = tid; idx < TS * TS / 4; idx += 256) {
    const int r = idx >> 5, cq = idx & 31;
    const float4 a = *reinterpret_cast<const float4*>(A + r * lda + cq * 4);
    As[r * PM_LDA + cq * 4 + 0] = a.x;
    As[r * PM_LDA + cq * 4 + 1] = a.y;
    As[r * PM_LDA + cq * 4 + 2] = a.z;
    As[r * PM_LDA + cq * 4 + 3] = a.w;
    *reinterpret_cast<float4*>(Bs + r * TS + cq * 4) = *reinterpret_cast<const float4*>(B + r * ldb + cq * 4);
  }
  __syncthreads();
  const int ty = tid >> 4, tx = tid & 15;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  for (int k = 0; k < TS; ++k) {
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = As[(ty * 8 + i) * PM_LDA + k];
    const float4 b0 = *reinterpret_cast<const float4*>(Bs + k * TS + tx * 8);
    const float4 b1 = *reinterpret_cast<const float4*>(Bs + k * TS + tx * 8 + 4);
    const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float* o = out + (ty * 8 + i) * TS + tx * 8;
    *reinterpret_cast<float4*>(o) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    *reinterpret_cast<float4*>(o + 4) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
  }
}

// Chained solve with pre-multiplied tiles.  Tile sequence of CTA c: the dependency tiles 0 .. c-2, the
// inverted diagonal block, W[blk]; a two-slot ring as in trsm_chain_kernel.
template <int DP, bool TRANS, bool F16>
__global__ void __launch_bounds__(256)
trsm_chain_w_kernel(const TrsmJobs jobs, int nblk, int* __restrict__ sync) {
  constexpr int NB8 = (DP + 7) / 8;
  constexpr int LT = trsm_ldt(TRANS);
  constexpr int LX = trsm_ldx(NB8);
  extern __shared__ __align__(16) float tsm[];
  float* ring = tsm;                  // [2][128][LT]
  float* xring = tsm + 2 * TS * LT;   // [2][128][LX]
  __half* xh = reinterpret_cast<__half*>(xring + 2 * TS * LX);   // [NB8*8][XH_LD] converted x (F16 only)
  __half* xl = xh + NB8 * 8 * XH_LD;
  __shared__ int s_ticket, s_next;
  __shared__ float s_red[8];
  if (threadIdx.x == 0) s_ticket = atomicAdd(sync, 1);
  __syncthreads();
  const int jid = s_ticket % jobs.njobs;
  const int c = s_ticket / jobs.njobs;
  const TrsmJob& J = jobs.job[jid];
  const float* __restrict__ L = J.L;
  const long long ldl = J.ldl;
  float* __restrict__ Bm = J.x;
  const int blk = TRANS ? nblk - 1 - c : c;
  int* flags = sync + 64 + jid * nblk;
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int m0 = (threadIdx.x >> 5) * 16;
  const int nt = c == 0 ? 1 : c + 1;                       // tiles in this CTA's sequence
  enum { DEP = 0, INV = 1, WT = 2 };
  auto kind_of = [&](int i) { return c == 0 ? INV : (i <= c - 2 ? DEP : (i == c - 1 ? INV : WT)); };
  auto src_of = [&](int d) { return TRANS ? nblk - 1 - d : d; };
  auto xblock_of = [&](int i) { return kind_of(i) == DEP ? src_of(i) : src_of(c - 1); };  // DEP / WT only
  auto issue_tile = [&](int i) {
    float* dst = ring + (i & 1) * TS * LT;
    const int kd = kind_of(i);
    if (kd == DEP) {
      const int src = src_of(i);
      const float* tile = TRANS ? L + static_cast<long long>(src) * TS * ldl + static_cast<long long>(blk) * TS
                                : L + static_cast<long long>(blk) * TS * ldl + static_cast<long long>(src) * TS;
      load_tile_128_async<TRANS>(tile, ldl, dst);
    } else if (kd == INV) {
      load_tile_128_async<TRANS>(J.invdiag + static_cast<long long>(blk) * TS * TS, TS, dst);
    } else {
      load_tile_128_async<TRANS>(J.W + (static_cast<long long>(blk) * 2 + (TRANS ? 1 : 0)) * TS * TS, TS, dst);
    }
  };
  auto issue_x = [&](int i) {
    const float* xsrc = Bm + static_cast<long long>(xblock_of(i)) * TS * DP;
    float* xs = xring + (i & 1) * TS * LX;
    for (int idx = threadIdx.x; idx < TS * (DP / 4); idx += 256)
      cp_async_16(xs + (idx / (DP / 4)) * LX + (idx % (DP / 4)) * 4, xsrc + idx * 4);
  };
  if (LX > DP)
    for (int idx = threadIdx.x; idx < 2 * TS * (LX - DP); idx += 256)
      xring[(idx / (LX - DP)) * LX + DP + idx % (LX - DP)] = 0.f;
  issue_tile(0);
  cp_async_commit();

  auto wait_flag = [&](const int* f, int what) {
    unsigned spins = 0;
    unsigned long long t0 = 0;
    while (ld_acquire_gpu(f) == 0) {
      if ((++spins & 0x3FFu) == 0) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        else if (now - t0 > 4000000000ull) {
          printf("sobolev_b200: trsm chain (w) timed out waiting for block %d (job %d cta %d)\n", what, jid, c);
          __trap();
        }
      }
    }
  };
  const long long frow = static_cast<long long>(blk) * TS + m0 + g;
  float* b0 = Bm + frow * DP;
  float* b1 = b0 + 8 * DP;
  float bfrag[NB8][4];
  auto load_rhs = [&]() {
#pragma unroll
    for (int nb = 0; nb < NB8; ++nb) {
      const int cc = nb * 8 + 2 * t;
      float2 v0 = make_float2(0.f, 0.f), v1 = make_float2(0.f, 0.f);
      if (cc < DP) {
        v0 = __ldcg(reinterpret_cast<const float2*>(J.rhs + frow * DP + cc));
        v1 = __ldcg(reinterpret_cast<const float2*>(J.rhs + (frow + 8) * DP + cc));
        if (J.V != nullptr) {
          const float2 a0 = *reinterpret_cast<const float2*>(J.V + frow * DP + cc);
          const float2 a1 = *reinterpret_cast<const float2*>(J.V + (frow + 8) * DP + cc);
          v0.x = fmaf(J.lambda, a0.x, v0.x); v0.y = fmaf(J.lambda, a0.y, v0.y);
          v1.x = fmaf(J.lambda, a1.x, v1.x); v1.y = fmaf(J.lambda, a1.y, v1.y);
        }
      }
      bfrag[nb][0] += v0.x; bfrag[nb][1] += v0.y; bfrag[nb][2] += v1.x; bfrag[nb][3] += v1.y;
    }
  };
#pragma unroll
  for (int nb = 0; nb < NB8; ++nb)
#pragma unroll
    for (int i = 0; i < 4; ++i) bfrag[nb][i] = 0.f;
  if (jid == 0) load_rhs();

  // acc = tile * x: fp16 hi/lo m16n8k16 (half the MMAs) or 3xTF32 m16n8k8
  auto product = [&](const float* Tsm, const float* xsm, int which, float (&acc)[NB8][4]) {
    if constexpr (F16) {
      const float ts = __ldg(J.scales + which);
      const float invx = convert_x<NB8>(xsm, xh, xl, s_red);
      tile_times_xh<NB8, TRANS>(Tsm, xh, xl, ts, acc);
      const float r = invx / ts;
#pragma unroll
      for (int nb = 0; nb < NB8; ++nb)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[nb][q] *= r;
    } else {
      tile_times_rhs<NB8, TRANS>(Tsm, xsm, acc);
    }
  };
  bool have_x = false;  // the x block of the current tile was prefetched by the previous iteration
  for (int i = 0; i < nt; ++i) {
    const int kd = kind_of(i);
    const bool need_x = kd != INV;
    if (threadIdx.x == 0) {
      if (need_x && !have_x) wait_flag(flags + xblock_of(i), xblock_of(i));
      if (kd == INV && jid == 1) wait_flag(sync + 64 + blk, blk);  // job 0 has published this block: the rhs
      int nx = 0;
      if (i + 1 < nt && kind_of(i + 1) != INV) nx = ld_acquire_gpu(flags + xblock_of(i + 1));
      s_next = nx;
    }
    __syncthreads();  // flags seen; everyone is done with slot (i + 1) & 1 of both rings
    const bool nxt = s_next != 0;
    if (need_x && !have_x) {
      issue_x(i);
      cp_async_commit();
    }
    if (i + 1 < nt) {
      issue_tile(i + 1);
      if (nxt) issue_x(i + 1);
      cp_async_commit();
      cp_async_wait<1>();  // everything but that newest group: tile i (and its x) have landed
    } else {
      cp_async_wait<0>();
    }
    if (kd == INV) {
      if (jid == 1) load_rhs();
      __syncthreads();  // tile landed for everyone (xring slot i & 1 carries no x for this tile: free)
      float* xs = xring + (i & 1) * TS * LX;
#pragma unroll
      for (int nb = 0; nb < NB8; ++nb) {
        const int cc = nb * 8 + 2 * t;
        xs[(m0 + g) * LX + cc] = bfrag[nb][0];
        xs[(m0 + g) * LX + cc + 1] = bfrag[nb][1];
        xs[(m0 + g + 8) * LX + cc] = bfrag[nb][2];
        xs[(m0 + g + 8) * LX + cc + 1] = bfrag[nb][3];
      }
      __syncthreads();
      float acc[NB8][4];
      product(ring + (i & 1) * TS * LT, xs, 1, acc);
#pragma unroll
      for (int nb = 0; nb < NB8; ++nb)
#pragma unroll
        for (int q = 0; q < 4; ++q) bfrag[nb][q] = acc[nb][q];
    } else {
      __syncthreads();
      float acc[NB8][4];
      product(ring + (i & 1) * TS * LT, xring + (i & 1) * TS * LX, kd == DEP ? 0 : 2, acc);
#pragma unroll
      for (int nb = 0; nb < NB8; ++nb)
#pragma unroll
        for (int q = 0; q < 4; ++q) bfrag[nb][q] -= acc[nb][q];
    }
    have_x = nxt;
  }
#pragma unroll
  for (int nb = 0; nb < NB8; ++nb) {
    const int cc = nb * 8 + 2 * t;
    if (cc < DP) {
      *reinterpret_cast<float2*>(b0 + cc) = make_float2(bfrag[nb][0], bfrag[nb][1]);
      *reinterpret_cast<float2*>(b1 + cc) = make_float2(bfrag[nb][2], bfrag[nb][3]);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    st_release_gpu(flags + blk, 1);
  }
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

// scales[i] = 2^(12 - e_i) with max_i = f 2^e_i the largest magnitude found by absmax (bits at scales[8 + i])
__global__ void trsm_scales_kernel(float* scales) {
  const int i = threadIdx.x;
  if (i < 3) {
    const float m = __uint_as_float(reinterpret_cast<unsigned*>(scales)[8 + i]);
    int e = 0;
    if (m > 0.f && isfinite(m)) frexpf(m, &e);
    scales[i] = ldexpf(1.f, 12 - e);
  }
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

__global__ void axpby_kernel(long long len, float a, const float* __restrict__ X, float b,
                             float* __restrict__ Y) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < len; i += stride)
    Y[i] = (b == 0.f) ? a * X[i] : fmaf(a, X[i], b * Y[i]);
}

// blockDim = (dp, rows_per_block): thread (c, ry) walks rows ry, ry + stride, ... of column c, so
// a warp reads consecutive floats of the row-major M x dp array.
constexpr int RED_THREADS = 512;

__device__ __forceinline__ void block_column_reduce(float v, int dp, float* out) {
  __shared__ float red[RED_THREADS];
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  red[tid] = v;
  __syncthreads();
  if (threadIdx.y == 0) {
    float s = 0.f;
    for (int r = 0; r < static_cast<int>(blockDim.y); ++r) s += red[r * dp + threadIdx.x];
    atomicAdd(out + threadIdx.x, s);
  }
}

__global__ void coldot_kernel(const float* __restrict__ P, const float* __restrict__ Q, long long M,
                              int dp, float* __restrict__ out) {
  float s = 0.f;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.y;
  for (long long r = static_cast<long long>(blockIdx.x) * blockDim.y + threadIdx.y; r < M; r += stride)
    s = fmaf(P[r * dp + threadIdx.x], Q[r * dp + threadIdx.x], s);
  block_column_reduce(s, dp, out);
}

__global__ void cg_update_kernel(long long M, int dp, const float* __restrict__ P,
                                 const float* __restrict__ AP, float* __restrict__ X,
                                 float* __restrict__ R, const float* __restrict__ rs_old,
                                 const float* __restrict__ pAp, float eps, float* __restrict__ rs_new) {
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
  if (AP != nullptr) block_column_reduce(s, dp, rs_new);
}

__global__ void cg_direction_kernel(long long M, int dp, const float* __restrict__ R,
                                    float* __restrict__ P, const float* __restrict__ rs_new,
                                    const float* __restrict__ rs_old) {
  const int c = threadIdx.x;
  const float ro = rs_old[c];
  const float beta = ro > 0.f ? rs_new[c] / ro : 0.f;  // padded / exactly converged columns
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.y;
  for (long long r = static_cast<long long>(blockIdx.x) * blockDim.y + threadIdx.y; r < M; r += stride) {
    const long long i = r * dp + c;
    P[i] = fmaf(beta, P[i], R[i]);
  }
}

static int red_grid(long long M, int rows_per_block) {
  long long g = ceil_div(M, static_cast<long long>(rows_per_block) * 4);
  if (g < 1) g = 1;
  if (g > 592) g = 592;
  return static_cast<int>(g);
}

template <int DP>
static int trsm_chain_launch(cudaStream_t st, const TrsmJobs& jobs, long long M, int transpose, int* sync) {
  const int nblk = static_cast<int>(M / TS);
  // workspace: [0] ticket counter (64 ints reserved), then one flag per block and job
  const int smem = (2 * TS * trsm_ldt(transpose != 0) + 2 * TS * trsm_ldx((DP + 7) / 8)) * static_cast<int>(sizeof(float));
  SAB_CHECK_CUDA(cudaMemsetAsync(sync, 0, (64 + 2 * nblk) * sizeof(int), st));
  bool premul = true;
  for (int j = 0; j < jobs.njobs; ++j) premul = premul && jobs.job[j].W != nullptr;
  if (premul) {
    // fp16 hi/lo tile products (half the MMAs) measured only 10 % faster than 3xTF32 -- the conversion of x and
    // the register split of the tile eat the saving -- so the TF32 form stays the default
    const char* tv = std::getenv("SAB_TRSM_F16");
    const bool f16 = tv && std::atoi(tv) != 0;
    const int smem_w = smem + (f16 ? 2 * ((DP + 7) / 8) * 8 * XH_LD * static_cast<int>(sizeof(__half)) : 0);
    auto go = [&](auto kern) -> int {
      SAB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_w));
      kern<<<nblk * jobs.njobs, 256, smem_w, st>>>(jobs, nblk, sync);
      return 0;
    };
    int rc;
    if (transpose) rc = f16 ? go(trsm_chain_w_kernel<DP, true, true>) : go(trsm_chain_w_kernel<DP, true, false>);
    else rc = f16 ? go(trsm_chain_w_kernel<DP, false, true>) : go(trsm_chain_w_kernel<DP, false, false>);
    if (rc) return rc;
  } else if (transpose) {
    auto kern = trsm_chain_kernel<DP, true>;
    SAB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kern<<<nblk * jobs.njobs, 256, smem, st>>>(jobs, nblk, sync);
  } else {
    auto kern = trsm_chain_kernel<DP, false>;
    SAB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kern<<<nblk * jobs.njobs, 256, smem, st>>>(jobs, nblk, sync);
  }
  SAB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <int DP>
static int trsm_launch(cudaStream_t st, const float* L, long long M, long long ldl, const float* invdiag,
                       float* B, int transpose, int* sync, const float* W = nullptr) {
  const int nblk = static_cast<int>(M / TS);
  const char* ev = std::getenv("SAB_TRSM_STEPS");
  if (sync != nullptr && !(ev && std::atoi(ev) != 0)) {
    TrsmJobs jobs = {};
    jobs.njobs = 1;
    jobs.job[0] = TrsmJob{L, ldl, invdiag, B, B, nullptr, 0.f, W, W ? W + (M / TS) * 2ll * TS * TS : nullptr};
    {
      const char* tp = std::getenv("SAB_TRSM_TRACE_PTR");
      jobs.trace = tp ? reinterpret_cast<long long*>(std::strtoull(tp, nullptr, 0)) : nullptr;
    }
    return trsm_chain_launch<DP>(st, jobs, M, transpose, sync);
  }
  const int smem = (TS * trsm_ldt(transpose != 0) + TS * trsm_ldx((DP + 7) / 8)) * static_cast<int>(sizeof(float));
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (transpose) {
    auto kern = trsm_step_kernel<DP, true>;
    SAB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    for (int k = nblk - 1; k >= 0; --k) {
      cfg.gridDim = dim3(k + 1);
      SAB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, L, ldl, invdiag, B, k, nblk));
    }
  } else {
    auto kern = trsm_step_kernel<DP, false>;
    SAB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    for (int k = 0; k < nblk; ++k) {
      cfg.gridDim = dim3(nblk - k);
      SAB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, L, ldl, invdiag, B, k, nblk));
    }
  }
  SAB_CHECK_CUDA(cudaGetLastError());
  return 0;
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
  const char* v = std::getenv("SAB_LINALG_SIMT");
  return !(v && std::atoi(v) != 0);
}

// width (in 128-blocks) of the outer panels of the two-level factorisation: the trailing matrix is
// read and written once per outer panel, so wider panels move the update from HBM-bound to tensor-bound
static int outer_blocks() {
  const char* v = std::getenv("SAB_CHOL_NB");
  int nb = v ? std::atoi(v) : 512;
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

extern "C" int sa_potrf(void* stream, float* A, int64_t M, int64_t lda, float* invdiag, int* info,
                        void* workspace, int64_t workspace_bytes) {
  SAB_REQUIRE_MAT(M, lda, A);
  SAB_REQUIRE(invdiag != nullptr && info != nullptr, "invdiag / info must not be NULL");
  SAB_REQUIRE_WS(M, workspace, workspace_bytes);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nblk = static_cast<int>(M / TS);
  const int dsmem = (TS * LDD + 4 * TS) * static_cast<int>(sizeof(float));
  SAB_CHECK_CUDA(cudaFuncSetAttribute(diag_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dsmem));
  SAB_CHECK_CUDA(cudaMemsetAsync(info, 0, sizeof(int), st));
  const bool tensor = use_tensor_path();
  const int nbb = tensor ? outer_blocks() : nblk;  // SIMT: classic right-looking (one "outer panel")
  const LinalgWs w = carve(workspace, M);
  if (tensor) {
    // |L_ij| <= sqrt(max_i A_ii): one power-of-two operand scale for the whole factorisation
    SAB_CHECK_CUDA(cudaMemsetAsync(w.hdr, 0, HDR_FLOATS * 4, st));
    absmax_kernel<<<static_cast<int>(ceil_div(M, 256)), 256, 0, st>>>(A, M, lda, 1,
                                                                      reinterpret_cast<unsigned*>(w.hdr));
    make_scale_kernel<<<1, 1, 0, st>>>(w.hdr, 1);
  }
  for (int k0 = 0; k0 < nblk; k0 += nbb) {
    const int kend = k0 + nbb < nblk ? k0 + nbb : nblk;
    // ---- factor the outer panel (columns [k0, kend) x all rows below) with 128-wide fp32 SIMT steps
    for (int k = k0; k < kend; ++k) {
      diag_factor_kernel<<<1, 256, dsmem, st>>>(A, lda, k, invdiag, info);
      const int rem = nblk - k - 1;
      if (rem == 0) break;
      sgemm_kernel<MODE_PANEL><<<rem, 256, 0, st>>>(A, lda, invdiag + static_cast<long long>(k) * TS * TS, TS, k,
                                                    nblk, 0.f, 0.f);
      const int jlim = (tensor ? kend : nblk) - k - 1;  // block columns right of k still inside the panel
      if (jlim >= rem)
        sgemm_kernel<MODE_SYRK><<<rem * (rem + 1) / 2, 256, 0, st>>>(A, lda, nullptr, 0, k, nblk,
                                                                     static_cast<float>(jlim), 0.f);
      else if (jlim > 0)
        sgemm_kernel<MODE_SYRK><<<rem * jlim, 256, 0, st>>>(A, lda, nullptr, 0, k, nblk, static_cast<float>(jlim), 0.f);
    }
    // ---- trailing update C -= P P^T on the tensor cores, P = the finished panel below row kend*128
    if (tensor && kend < nblk) {
      const long long rows = static_cast<long long>(nblk - kend) * TS;
      const int W = (kend - k0) * TS;
      float* panel = A + static_cast<long long>(kend) * TS * lda + static_cast<long long>(k0) * TS;
      float* C = A + static_cast<long long>(kend) * TS * (lda + 1);
      exact_diag_kernel<<<static_cast<int>(ceil_div(rows * 32, 256)), 256, 0, st>>>(C, lda, panel, lda, rows, W, -1.f,
                                                                                w.dsave);
      split_f16_kernel<<<static_cast<int>(ceil_div(rows * (W / 8), 256)), 256, 0, st>>>(panel, lda, rows, W, 0, 0, w.hdr,
                                                                                     w.opA, w.opB);
      SAB_CHECK_CUDA(cudaGetLastError());
      if (int rc = tc_gemm_nt(stream, w.opA, rows, 3 * W, w.opB, rows, 3 * W, 3 * W, C, lda, -tc_gain(W), w.hdr + 2,
                              1.f, 1))
        return rc;
      restore_diag_kernel<<<static_cast<int>(ceil_div(rows, 256)), 256, 0, st>>>(C, lda, rows, w.dsave);
    }
  }
  SAB_CHECK_CUDA(cudaGetLastError());
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

extern "C" int64_t sa_trsm_workspace_bytes(int64_t M) { return M <= 0 ? -1 : (64 + 2 * (M / TS) + 2) * 4; }

template <int DP>
static int trsm_pair_dp(cudaStream_t st, const TrsmJobs& jobs, long long M, int transpose, int* sync) {
  return trsm_chain_launch<DP>(st, jobs, M, transpose, sync);
}

extern "C" int64_t sa_trsm_premul_floats(int64_t M) {
  return M <= 0 ? -1 : (M / TS) * 2 * static_cast<int64_t>(TS) * TS + 64;
}

extern "C" int sa_trsm_premul(void* stream, const float* L, int64_t M, int64_t ldl, const float* invdiag,
                              float* W) {
  SAB_REQUIRE_MAT(M, ldl, L);
  SAB_REQUIRE(invdiag != nullptr && W != nullptr && (reinterpret_cast<uintptr_t>(W) & 15) == 0, "bad operands");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nblk = static_cast<int>(M / TS);
  const int smem = (TS * PM_LDA + TS * TS) * static_cast<int>(sizeof(float));
  SAB_CHECK_CUDA(cudaFuncSetAttribute(trsm_premul_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  trsm_premul_kernel<<<dim3(nblk, 2), 256, smem, st>>>(L, ldl, invdiag, nblk, W);
  // power-of-two scales of the fp16 tile products behind the tiles: {L, inverted diagonal blocks, W}
  const long long tiles = static_cast<long long>(nblk) * 2 * TS * TS;
  float* scales = W + tiles;
  SAB_CHECK_CUDA(cudaMemsetAsync(scales, 0, 64 * sizeof(float), st));
  unsigned* bits = reinterpret_cast<unsigned*>(scales) + 8;
  absmax_kernel<<<1184, 256, 0, st>>>(L, M, ldl, 0, bits + 0);
  absmax_kernel<<<static_cast<int>(ceil_div(static_cast<long long>(nblk) * TS, 8)), 256, 0, st>>>(
      invdiag, static_cast<long long>(nblk) * TS, TS, 2, bits + 1);
  absmax_kernel<<<static_cast<int>(ceil_div(2ll * nblk * TS, 8)), 256, 0, st>>>(W, 2ll * nblk * TS, TS, 2, bits + 2);
  trsm_scales_kernel<<<1, 32, 0, st>>>(scales);
  SAB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int sa_trsm_pair_w(void* stream, const float* La, int64_t lda, const float* inva, const float* Wa,
                              const float* Lb, int64_t ldb, const float* invb, const float* Wb, int64_t M,
                              float* Xa, float* Xb, const float* V, float lambda, int dp, int transpose,
                              void* workspace, int64_t workspace_bytes) {
  SAB_REQUIRE_MAT(M, lda, La);
  SAB_REQUIRE_MAT(M, ldb, Lb);
  SAB_REQUIRE(inva != nullptr && invb != nullptr && Xa != nullptr && Xb != nullptr && Xa != Xb, "bad operands");
  SAB_REQUIRE(workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0 &&
                  workspace_bytes >= sa_trsm_workspace_bytes(M),
              "workspace must hold sa_trsm_workspace_bytes(M) = %lld bytes", (long long)sa_trsm_workspace_bytes(M));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TrsmJobs jobs = {};
  jobs.njobs = 2;
  const long long tiles = (M / TS) * 2ll * TS * TS;
  jobs.job[0] = TrsmJob{La, lda, inva, Xa, Xa, nullptr, 0.f, Wa, Wa ? Wa + tiles : nullptr};
  jobs.job[1] = TrsmJob{Lb, ldb, invb, Xa, Xb, V, lambda, Wb, Wb ? Wb + tiles : nullptr};
  int* sync = static_cast<int*>(workspace);
  switch (dp) {
    case 4: return trsm_pair_dp<4>(st, jobs, M, transpose, sync);
    case 8: return trsm_pair_dp<8>(st, jobs, M, transpose, sync);
    case 12: return trsm_pair_dp<12>(st, jobs, M, transpose, sync);
    case 16: return trsm_pair_dp<16>(st, jobs, M, transpose, sync);
    case 20: return trsm_pair_dp<20>(st, jobs, M, transpose, sync);
    case 24: return trsm_pair_dp<24>(st, jobs, M, transpose, sync);
    case 28: return trsm_pair_dp<28>(st, jobs, M, transpose, sync);
    case 32: return trsm_pair_dp<32>(st, jobs, M, transpose, sync);
  }
  set_error("dp must be a multiple of 4 in [4, 32], got %d", dp);
  return 2;
}

extern "C" int sa_trsm_pair(void* stream, const float* La, int64_t lda, const float* inva, const float* Lb,
                            int64_t ldb, const float* invb, int64_t M, float* Xa, float* Xb, const float* V,
                            float lambda, int dp, int transpose, void* workspace, int64_t workspace_bytes) {
  return sa_trsm_pair_w(stream, La, lda, inva, nullptr, Lb, ldb, invb, nullptr, M, Xa, Xb, V, lambda, dp, transpose,
                        workspace, workspace_bytes);
}

extern "C" int sa_trsm(void* stream, const float* L, int64_t M, int64_t ldl, const float* invdiag,
                       float* B, int dp, int transpose, void* workspace, int64_t workspace_bytes) {
  SAB_REQUIRE_MAT(M, ldl, L);
  SAB_REQUIRE(invdiag != nullptr && B != nullptr, "NULL operand");
  SAB_REQUIRE(workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0 &&
                  workspace_bytes >= sa_trsm_workspace_bytes(M),
              "workspace must hold sa_trsm_workspace_bytes(M) = %lld bytes", (long long)sa_trsm_workspace_bytes(M));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int* sync = static_cast<int*>(workspace);
  switch (dp) {
    case 4: return trsm_launch<4>(st, L, M, ldl, invdiag, B, transpose, sync);
    case 8: return trsm_launch<8>(st, L, M, ldl, invdiag, B, transpose, sync);
    case 12: return trsm_launch<12>(st, L, M, ldl, invdiag, B, transpose, sync);
    case 16: return trsm_launch<16>(st, L, M, ldl, invdiag, B, transpose, sync);
    case 20: return trsm_launch<20>(st, L, M, ldl, invdiag, B, transpose, sync);
    case 24: return trsm_launch<24>(st, L, M, ldl, invdiag, B, transpose, sync);
    case 28: return trsm_launch<28>(st, L, M, ldl, invdiag, B, transpose, sync);
    case 32: return trsm_launch<32>(st, L, M, ldl, invdiag, B, transpose, sync);
  }
  set_error("dp must be a multiple of 4 in [4, 32], got %d", dp);
  return 2;
}

#define SAB_REQUIRE_VEC(M, dp) \
  SAB_REQUIRE((M) > 0 && (dp) >= 1 && (dp) <= 32, "bad vector shape M=%lld dp=%d", (long long)(M), (dp))

extern "C" int sa_coldot(void* stream, const float* P, const float* Q, int64_t M, int dp, float* out) {
  SAB_REQUIRE_VEC(M, dp);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SAB_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * dp, st));
  const int rows = RED_THREADS / dp;
  coldot_kernel<<<red_grid(M, rows), dim3(dp, rows), 0, st>>>(P, Q, M, dp, out);
  SAB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int sa_cg_update(void* stream, int64_t M, int dp, const float* P, const float* AP, float* X,
                            float* R, const float* rs_old, const float* pAp, float eps, float* rs_new) {
  SAB_REQUIRE_VEC(M, dp);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (AP != nullptr) SAB_CHECK_CUDA(cudaMemsetAsync(rs_new, 0, sizeof(float) * dp, st));
  const int rows = RED_THREADS / dp;
  cg_update_kernel<<<red_grid(M, rows), dim3(dp, rows), 0, st>>>(M, dp, P, AP, X, R, rs_old, pAp, eps, rs_new);
  SAB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int sa_cg_direction(void* stream, int64_t M, int dp, const float* R, float* P,
                               const float* rs_new, const float* rs_old) {
  SAB_REQUIRE_VEC(M, dp);
  const int rows = RED_THREADS / dp;
  cg_direction_kernel<<<red_grid(M, rows), dim3(dp, rows), 0, static_cast<cudaStream_t>(stream)>>>(
      M, dp, R, P, rs_new, rs_old);
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
