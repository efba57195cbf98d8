// Fused kernel-matrix x matrix product for sm_100a:   out[a x dp] += k(A, B) @ V      (sa_kmv)
// and the dense variant                               out[a x b]   = k(A, B)          (sa_kdense)
//
// One persistent CTA per SM, warp-specialised:
//   warp 0      TMA producer: 128x64 tile of A and 256x64 tile of B (fp16, 128B swizzle) per stage
//   warp 1      MMA issuer: tcgen05.mma kind::f16, M=128 N=256 K=16, fp32 accumulator in TMEM
//               (two 256-column accumulators = all 512 TMEM columns, double buffered)
//   warp 2      TMEM allocator + bulk loader of the V tile [256 x dp] and the B-norm tile
//   warps 4-11  epilogue: tcgen05.ld the dot products, d2 = |a|^2 + |b|^2 - 2 a.b, kernel function
//               (sqrt / exp2 on the SFU), then the thin product with V on the FMA pipe; each thread
//               owns one row of the tile and dp running sums, flushed with red.global.add when the
//               work unit ends.  K(A, B) only ever exists in TMEM and registers.
//
// Work decomposition: units = row tiles x nsplit column chunks, round-robin over the CTAs with the
// chunk index varying fastest: the CTAs running at the same time then cover only sms/nsplit row
// tiles, so their A strips (128 x p fp16 each, re-read once per column tile) stay L2 resident while
// CTAs on the same chunk sweep the same B tiles in step.
//
// Replaces falkon `Kernel.mmv / dmmv / __call__` [external dependency of the reference], reached
// from /root/reference/sobolev_alignment/krr_approx.py:227,314 and sobolev_alignment.py:950,955-958.
#include <cuda.h>

#include <cstdlib>
#include <mutex>

#include "../../include/sobolev_b200.h"
#include "common.cuh"
#include "ptx.cuh"

namespace sab {
namespace kmv {

constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BK = 64;
constexpr int UMMA_K = 16;
constexpr int A_BYTES = BM * BK * 2;
constexpr int B_BYTES = BN * BK * 2;
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int FIRST_EPI_WARP = 4;
constexpr int EPI_WARPS = 8;
constexpr int THREADS = (FIRST_EPI_WARP + EPI_WARPS) * 32;
constexpr int MAX_STAGES = 4;
constexpr int TMEM_COLS = 512;
constexpr int NUM_BARS = 16;
constexpr float LOG2E = 1.4426950408889634f;
// The tcgen05 fp32 accumulator is not round-to-nearest: every K=16 MMA step truncates ~2^-22.4 of the
// running sum.  For rows that coincide (every Nystrom centre is also a training row) the products
// are all positive, the dot product comes out low by ACC_BIAS_PER_STEP * steps (measured on B200 with
// tools/gpu_check.py accum: 0.63e-7 .. 0.85e-7 per step for 32..256 steps, +-15 % between rows), and
// sqrt() turns that into a visible error of k(x, x).  The epilogue therefore (1) rescales the dot
// product by the calibrated gain -- exact to first order for coincident / nearly coincident rows, a
// no-op for the typical pair whose centred dot product is ~0 -- and (2) snaps what is left below the
// residual resolution to exactly zero, so that k(x, x) = 1.
constexpr float ACC_BIAS_PER_STEP = 0.8e-7f;
constexpr float SNAP_PER_STEP = 0.3e-7f;
constexpr float SNAP_FLOOR = 4.0e-7f;

struct Params {
  long long a, b;   // rows of A / rows of B
  int kblocks;      // p_pad / 64
  int ta, tb;       // row tiles, column tiles
  int nsplit;       // column chunks per row tile
  int units;        // ta * nsplit
  int stages, vbufs;
  const float* nA;
  const float* nB;
  const float* V;
  float* out;
  long long ldo;
  float c0;         // gaussian: kscale*log2e ; others: kscale * sqrt(nu factor)
  float snap;       // squared distances below snap * (|a|^2 + |b|^2) are treated as exactly zero
  float m2g;        // -2 * (1 + accumulator-bias compensation)
  int symmetric;
};

template <int KID>
__device__ __forceinline__ float kernel_value(float d2, float c0) {
  if (KID == SA_KERNEL_GAUSSIAN) {
    return ex2_approx(-c0 * d2);
  } else {
    const float s = c0 * sqrt_approx(d2);  // laplacian: r ; matern: sqrt(2 nu) r
    const float e = ex2_approx(-LOG2E * s);
    if (KID == SA_KERNEL_LAPLACIAN) return e;
    if (KID == SA_KERNEL_MATERN15) return fmaf(s, e, e);
    return fmaf(fmaf(s, 0.33333333333f, 1.0f), s, 1.0f) * e;
  }
}

template <int KID, int DP, bool DENSE>
__global__ void __launch_bounds__(THREADS, 1)
kmv_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
           const Params P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  constexpr int VBUF_FLOATS = DENSE ? 0 : BN * DP;
  uint8_t* tiles = smem;
  float* vbuf = reinterpret_cast<float*>(tiles + P.stages * STAGE_BYTES);
  float* nbuf = vbuf + P.vbufs * VBUF_FLOATS;
  uint64_t* bars = reinterpret_cast<uint64_t*>(nbuf + P.vbufs * BN);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NUM_BARS);

  const uint32_t bar_full = smem_u32(bars + 0);     // [MAX_STAGES] TMA -> MMA
  const uint32_t bar_empty = smem_u32(bars + 4);    // [MAX_STAGES] MMA -> TMA
  const uint32_t bar_tfull = smem_u32(bars + 8);    // [2] MMA -> epilogue (accumulator ready)
  const uint32_t bar_tempty = smem_u32(bars + 10);  // [2] epilogue -> MMA (accumulator drained)
  const uint32_t bar_vfull = smem_u32(bars + 12);   // [2] V / norm tile landed
  const uint32_t bar_vempty = smem_u32(bars + 14);  // [2] V / norm tile consumed

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < MAX_STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_tfull + 8 * i, 1);
      mbar_init(bar_tempty + 8 * i, EPI_WARPS);
      mbar_init(bar_vfull + 8 * i, 1);
      mbar_init(bar_vempty + 8 * i, EPI_WARPS);
    }
    fence_barrier_init();
    prefetch_tensormap(&tmA);
    prefetch_tensormap(&tmB);
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int unit = blockIdx.x; unit < P.units; unit += gridDim.x) {
        const int rt = unit / P.nsplit;
        const int ch = unit % P.nsplit;
        const int j0 = static_cast<int>(static_cast<long long>(ch) * P.tb / P.nsplit);
        const int j1 = static_cast<int>(static_cast<long long>(ch + 1) * P.tb / P.nsplit);
        for (int j = j0; j < j1; ++j) {
          for (int kb = 0; kb < P.kblocks; ++kb) {
            mbar_wait(bar_empty + 8 * s, ph ^ 1);
            const uint32_t full = bar_full + 8 * s;
            mbar_arrive_expect_tx(full, STAGE_BYTES);
            const uint32_t dstA = smem_u32(tiles + s * STAGE_BYTES);
            tma_load_2d(dstA, &tmA, full, kb * BK, rt * BM);
            tma_load_2d(dstA + A_BYTES, &tmB, full, kb * BK, j * BN);
            if (++s == P.stages) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ====================================== MMA issuer ======================================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(BM, BN);
      int s = 0;
      uint32_t ph = 0;
      int buf = 0;
      uint32_t tph = 0;
      for (int unit = blockIdx.x; unit < P.units; unit += gridDim.x) {
        const int ch = unit % P.nsplit;
        const int j0 = static_cast<int>(static_cast<long long>(ch) * P.tb / P.nsplit);
        const int j1 = static_cast<int>(static_cast<long long>(ch + 1) * P.tb / P.nsplit);
        for (int j = j0; j < j1; ++j) {
          mbar_wait(bar_tempty + 8 * buf, ((tph >> buf) & 1u) ^ 1u);
          tcgen05_fence_after();
          const uint32_t tacc = tmem_base + static_cast<uint32_t>(buf * BN);
          for (int kb = 0; kb < P.kblocks; ++kb) {
            mbar_wait(bar_full + 8 * s, ph);
            tcgen05_fence_after();
            const uint32_t sa = smem_u32(tiles + s * STAGE_BYTES);
            const uint64_t da = umma_smem_desc_sw128(sa);
            const uint64_t db = umma_smem_desc_sw128(sa + A_BYTES);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              // advance 16 fp16 = 32 B along K inside the 128 B swizzle atom: +2 in 16-byte units
              umma_f16_ss(tacc, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k),
                          idesc, (kb | k) != 0 ? 1u : 0u);
            }
            umma_commit(bar_empty + 8 * s);  // stage reusable once these MMAs retire
            if (kb == P.kblocks - 1) umma_commit(bar_tfull + 8 * buf);
            if (++s == P.stages) { s = 0; ph ^= 1; }
          }
          tph ^= (1u << buf);
          buf ^= 1;
        }
      }
    }
  } else if (warp == 2) {
    // ================================ V / norm tile loader ==================================
    if (lane == 0) {
      int vb = 0;
      uint32_t vph = 0;
      for (int unit = blockIdx.x; unit < P.units; unit += gridDim.x) {
        const int ch = unit % P.nsplit;
        const int j0 = static_cast<int>(static_cast<long long>(ch) * P.tb / P.nsplit);
        const int j1 = static_cast<int>(static_cast<long long>(ch + 1) * P.tb / P.nsplit);
        for (int j = j0; j < j1; ++j) {
          mbar_wait(bar_vempty + 8 * vb, vph ^ 1);
          const long long col0 = static_cast<long long>(j) * BN;
          const int cols_valid = static_cast<int>(min(static_cast<long long>(BN), P.b - col0));
          const uint32_t full = bar_vfull + 8 * vb;
          const uint32_t vbytes = DENSE ? 0u : static_cast<uint32_t>(cols_valid * DP * 4);
          mbar_arrive_expect_tx(full, BN * 4 + vbytes);
          bulk_load_1d(smem_u32(nbuf + vb * BN), P.nB + col0, BN * 4, full);
          if (!DENSE)
            bulk_load_1d(smem_u32(vbuf + vb * VBUF_FLOATS), P.V + col0 * DP, vbytes, full);
          if (++vb == P.vbufs) { vb = 0; vph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp >= FIRST_EPI_WARP) {
    // ======================================= epilogue =======================================
    const int e = warp - FIRST_EPI_WARP;
    const int quarter = warp & 3;  // TMEM lane quarter this warp may read
    const int half = e >> 2;       // which 128 columns of the 256-wide accumulator
    const int row_in_tile = quarter * 32 + lane;
    const uint32_t tlane = static_cast<uint32_t>(quarter * 32) << 16;
    int buf = 0;
    uint32_t tph = 0;
    int vb = 0;
    uint32_t vph = 0;
    for (int unit = blockIdx.x; unit < P.units; unit += gridDim.x) {
      const int rt = unit / P.nsplit;
      const int ch = unit % P.nsplit;
      const int j0 = static_cast<int>(static_cast<long long>(ch) * P.tb / P.nsplit);
      const int j1 = static_cast<int>(static_cast<long long>(ch + 1) * P.tb / P.nsplit);
      const long long grow = static_cast<long long>(rt) * BM + row_in_tile;
      const bool row_ok = grow < P.a;
      const float na = row_ok ? __ldg(P.nA + grow) : 0.f;
      const long long sym_row = P.symmetric ? grow : -1;
      float acc[DP];
#pragma unroll
      for (int c = 0; c < DP; ++c) acc[c] = 0.f;

      for (int j = j0; j < j1; ++j) {
        const long long col0 = static_cast<long long>(j) * BN;
        const int cols_valid = static_cast<int>(min(static_cast<long long>(BN), P.b - col0));
        mbar_wait(bar_vfull + 8 * vb, vph);
        mbar_wait(bar_tfull + 8 * buf, (tph >> buf) & 1u);
        tcgen05_fence_after();
        const float* Vs = vbuf + vb * VBUF_FLOATS;
        const float* Ns = nbuf + vb * BN;
        const int cbeg = half * (BN / 2);
        const int cend = min(cbeg + BN / 2, cols_valid);
        const uint32_t tacc = tmem_base + tlane + static_cast<uint32_t>(buf * BN);
        for (int c0 = cbeg; c0 < cend; c0 += 32) {
          uint32_t r[32];
          tmem_ld_32x32(tacc + static_cast<uint32_t>(c0), r);
          tmem_wait_ld();
          const bool full_chunk = (c0 + 32 <= cend);
          if (DENSE) {
            float kv[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              const float nn = na + Ns[c0 + c];
              float d2 = fmaf(P.m2g, __uint_as_float(r[c]), nn);
              if (d2 < P.snap * nn || col0 + c0 + c == sym_row) d2 = 0.f;
              kv[c] = kernel_value<KID>(d2, P.c0);
            }
            if (row_ok) {
              float* dst = P.out + grow * P.ldo + col0 + c0;
              if (full_chunk && ((P.ldo & 3) == 0)) {
#pragma unroll
                for (int q = 0; q < 8; ++q)
                  reinterpret_cast<float4*>(dst)[q] =
                      make_float4(kv[4 * q], kv[4 * q + 1], kv[4 * q + 2], kv[4 * q + 3]);
              } else {
#pragma unroll
                for (int c = 0; c < 32; ++c)
                  if (c0 + c < cend) dst[c] = kv[c];
              }
            }
          } else {
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              if (full_chunk || (c0 + c < cend)) {
                const float nn = na + Ns[c0 + c];
                float d2 = fmaf(P.m2g, __uint_as_float(r[c]), nn);
                if (d2 < P.snap * nn || col0 + c0 + c == sym_row) d2 = 0.f;
                const float kv = kernel_value<KID>(d2, P.c0);
                const float4* vrow = reinterpret_cast<const float4*>(Vs + (c0 + c) * DP);
#pragma unroll
                for (int q = 0; q < DP / 4; ++q) {
                  const float4 v = vrow[q];
                  acc[4 * q + 0] = fmaf(kv, v.x, acc[4 * q + 0]);
                  acc[4 * q + 1] = fmaf(kv, v.y, acc[4 * q + 1]);
                  acc[4 * q + 2] = fmaf(kv, v.z, acc[4 * q + 2]);
                  acc[4 * q + 3] = fmaf(kv, v.w, acc[4 * q + 3]);
                }
              }
            }
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(bar_tempty + 8 * buf);
          mbar_arrive(bar_vempty + 8 * vb);
        }
        tph ^= (1u << buf);
        buf ^= 1;
        if (++vb == P.vbufs) { vb = 0; vph ^= 1; }
      }
      if (!DENSE && row_ok) {
        float* dst = P.out + grow * DP;
#pragma unroll
        for (int c = 0; c < DP; ++c) red_add_f32(dst + c, acc[c]);
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

static int make_tmap(CUtensorMap* tm, const void* base, long long rows, long long p_pad,
                     long long ld, int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  SAB_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(p_pad), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {BK, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstride,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SAB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
  return 0;
}

struct DeviceInfo {
  int sms = 0;
  int max_smem = 0;
  bool ok = false;
};

static int device_info(DeviceInfo* out) {
  static DeviceInfo cache[64];
  int dev = 0;
  SAB_CHECK_CUDA(cudaGetDevice(&dev));
  SAB_REQUIRE(dev >= 0 && dev < 64, "device ordinal %d out of range", dev);
  if (!cache[dev].ok) {
    int major = 0;
    SAB_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    SAB_REQUIRE(major == 10, "sobolev_b200 needs an sm_100 device (compute capability 10.x), found %d.x",
                major);
    SAB_CHECK_CUDA(cudaDeviceGetAttribute(&cache[dev].sms, cudaDevAttrMultiProcessorCount, dev));
    SAB_CHECK_CUDA(
        cudaDeviceGetAttribute(&cache[dev].max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    cache[dev].ok = true;
  }
  *out = cache[dev];
  return 0;
}

static int env_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return v ? std::atoi(v) : dflt;
}

// cost model: waves x (column tiles per unit + fixed per-unit overhead), subject to the A strips of
// the concurrently running row tiles fitting an L2 budget
static int choose_nsplit(int ta, int tb, int sms, long long p_pad) {
  int forced = env_int("SAB_NSPLIT", 0);
  if (forced > 0) return forced < tb ? forced : tb;
  const double l2_budget = static_cast<double>(env_int("SAB_L2_BUDGET_MB", 32)) * 1048576.0;
  const double strip = static_cast<double>(BM) * static_cast<double>(p_pad) * 2.0;
  const int live = ta < sms ? ta : sms;
  int min_ns = static_cast<int>(live * strip / l2_budget) + 1;
  if (min_ns > tb) min_ns = tb;
  if (min_ns < 1) min_ns = 1;
  double best = 1e300;
  int best_ns = min_ns;
  const int max_ns = tb < 1024 ? tb : 1024;
  for (int ns = min_ns; ns <= max_ns; ++ns) {
    const double waves = static_cast<double>(ceil_div(static_cast<int64_t>(ta) * ns, sms));
    const double len = static_cast<double>(ceil_div(tb, ns)) + 0.35;
    const double cost = waves * len;
    if (cost < best * 0.999) {
      best = cost;
      best_ns = ns;
    }
  }
  return best_ns;
}

template <int KID, int DP, bool DENSE>
static int launch_inst(cudaStream_t st, const CUtensorMap& tmA, const CUtensorMap& tmB,
                       const Params& P, int grid, int smem) {
  auto kern = kmv_kernel<KID, DP, DENSE>;
  SAB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  kern<<<grid, THREADS, smem, st>>>(tmA, tmB, P);
  SAB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <int KID>
static int launch_kid(cudaStream_t st, const CUtensorMap& tmA, const CUtensorMap& tmB,
                      const Params& P, int grid, int smem, int dp, bool dense) {
  if (dense) return launch_inst<KID, 4, true>(st, tmA, tmB, P, grid, smem);
  switch (dp) {
    case 4: return launch_inst<KID, 4, false>(st, tmA, tmB, P, grid, smem);
    case 8: return launch_inst<KID, 8, false>(st, tmA, tmB, P, grid, smem);
    case 12: return launch_inst<KID, 12, false>(st, tmA, tmB, P, grid, smem);
    case 16: return launch_inst<KID, 16, false>(st, tmA, tmB, P, grid, smem);
    case 20: return launch_inst<KID, 20, false>(st, tmA, tmB, P, grid, smem);
    case 24: return launch_inst<KID, 24, false>(st, tmA, tmB, P, grid, smem);
    case 28: return launch_inst<KID, 28, false>(st, tmA, tmB, P, grid, smem);
    case 32: return launch_inst<KID, 32, false>(st, tmA, tmB, P, grid, smem);
  }
  set_error("dp must be a multiple of 4 in [4, 32], got %d", dp);
  return 2;
}

int launch(void* stream, int kernel, float kscale, const void* A, const float* nA, int64_t a,
           int64_t lda, const void* B, const float* nB, int64_t b, int64_t ldb, int64_t p_pad,
           const float* V, int dp, float* out, int64_t ldo, int symmetric, bool dense) {
  SAB_REQUIRE(kernel >= 0 && kernel <= 3, "unknown kernel id %d", kernel);
  SAB_REQUIRE(a >= 0 && b >= 0, "negative row count");
  if (a == 0 || b == 0) return 0;
  SAB_REQUIRE(p_pad > 0 && p_pad % BK == 0, "p_pad (%lld) must be a positive multiple of %d",
              static_cast<long long>(p_pad), BK);
  SAB_REQUIRE(lda >= p_pad && ldb >= p_pad && lda % 8 == 0 && ldb % 8 == 0,
              "row strides must be >= p_pad and multiples of 8 elements (lda=%lld ldb=%lld)",
              static_cast<long long>(lda), static_cast<long long>(ldb));
  SAB_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(nB) & 15) == 0,
              "operands must be 16-byte aligned");
  SAB_REQUIRE(a < (1ll << 31) && b < (1ll << 31), "row counts must fit in int32");
  if (!dense) {
    SAB_REQUIRE(V != nullptr && (reinterpret_cast<uintptr_t>(V) & 15) == 0, "V must be 16-byte aligned");
    SAB_REQUIRE(dp >= 4 && dp <= 32 && dp % 4 == 0, "dp must be a multiple of 4 in [4, 32], got %d", dp);
  } else {
    SAB_REQUIRE(ldo >= b, "ldo (%lld) < b (%lld)", static_cast<long long>(ldo), static_cast<long long>(b));
    SAB_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "out must be 16-byte aligned");
  }
  DeviceInfo di;
  if (int rc = device_info(&di)) return rc;

  Params P;
  P.a = a;
  P.b = b;
  P.kblocks = static_cast<int>(p_pad / BK);
  P.ta = static_cast<int>(ceil_div(a, BM));
  P.tb = static_cast<int>(ceil_div(b, BN));
  P.nsplit = dense ? 1 : choose_nsplit(P.ta, P.tb, di.sms, p_pad);
  P.units = P.ta * P.nsplit;
  P.nA = nA;
  P.nB = nB;
  P.V = V;
  P.out = out;
  P.ldo = ldo;
  P.symmetric = symmetric;
  {
    const float steps = static_cast<float>(P.kblocks * (BK / UMMA_K));
    const char* sv = std::getenv("SAB_SNAP");   // < 0 disables both corrections (calibration runs)
    const float forced = sv ? static_cast<float>(std::atof(sv)) : 0.f;
    P.snap = sv ? forced : SNAP_PER_STEP * steps + SNAP_FLOOR;
    P.m2g = (sv && forced < 0.f) ? -2.f : -2.f * (1.f + ACC_BIAS_PER_STEP * steps);
  }
  if (kernel == SA_KERNEL_GAUSSIAN) P.c0 = kscale * LOG2E;
  else if (kernel == SA_KERNEL_LAPLACIAN) P.c0 = kscale;
  else if (kernel == SA_KERNEL_MATERN15) P.c0 = kscale * 1.7320508075688772f;
  else P.c0 = kscale * 2.23606797749979f;

  // shared memory plan: as many pipeline stages as fit, two V buffers when they fit too
  const int vbytes = (dense ? 0 : BN * dp * 4) + BN * 4;
  const int fixed = 1024 /*alignment slack*/ + NUM_BARS * 8 + 64;
  int stages = env_int("SAB_STAGES", MAX_STAGES);
  int vbufs = env_int("SAB_VBUFS", 2);
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages < 2) stages = 2;
  if (vbufs < 1) vbufs = 1;
  if (vbufs > 2) vbufs = 2;
  if (stages * STAGE_BYTES + vbufs * vbytes + fixed > di.max_smem) vbufs = 1;
  while (stages > 2 && stages * STAGE_BYTES + vbufs * vbytes + fixed > di.max_smem) --stages;
  const int smem = stages * STAGE_BYTES + vbufs * vbytes + fixed;
  SAB_REQUIRE(smem <= di.max_smem, "shared memory plan (%d B) exceeds the device limit (%d B)", smem,
              di.max_smem);
  P.stages = stages;
  P.vbufs = vbufs;

  CUtensorMap tmA, tmB;
  if (int rc = make_tmap(&tmA, A, a, p_pad, lda, BM)) return rc;
  if (int rc = make_tmap(&tmB, B, b, p_pad, ldb, BN)) return rc;

  const int grid = P.units < di.sms ? P.units : di.sms;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (kernel) {
    case SA_KERNEL_GAUSSIAN: return launch_kid<SA_KERNEL_GAUSSIAN>(st, tmA, tmB, P, grid, smem, dp, dense);
    case SA_KERNEL_LAPLACIAN: return launch_kid<SA_KERNEL_LAPLACIAN>(st, tmA, tmB, P, grid, smem, dp, dense);
    case SA_KERNEL_MATERN15: return launch_kid<SA_KERNEL_MATERN15>(st, tmA, tmB, P, grid, smem, dp, dense);
    default: return launch_kid<SA_KERNEL_MATERN25>(st, tmA, tmB, P, grid, smem, dp, dense);
  }
}

int query_device(int* sms, int* smem) {
  DeviceInfo di;
  if (int rc = device_info(&di)) return rc;
  if (sms) *sms = di.sms;
  if (smem) *smem = di.max_smem;
  return 0;
}

}  // namespace kmv
}  // namespace sab

extern "C" int sa_kmv(void* stream, int kernel, float kscale, const void* A, const float* nA,
                      int64_t a, int64_t lda, const void* B, const float* nB, int64_t b, int64_t ldb,
                      int64_t p_pad, const float* V, int dp, float* out, int symmetric) {
  return sab::kmv::launch(stream, kernel, kscale, A, nA, a, lda, B, nB, b, ldb, p_pad, V, dp, out, dp,
                          symmetric, false);
}

extern "C" int sa_kdense(void* stream, int kernel, float kscale, const void* A, const float* nA,
                         int64_t a, int64_t lda, const void* B, const float* nB, int64_t b,
                         int64_t ldb, int64_t p_pad, float* out, int64_t ldo, int symmetric) {
  return sab::kmv::launch(stream, kernel, kscale, A, nA, a, lda, B, nB, b, ldb, p_pad, nullptr, 4, out,
                          ldo, symmetric, true);
}

extern "C" int sa_device_info(int* sm_count, int* smem_bytes) {
  return sab::kmv::query_device(sm_count, smem_bytes);
}
