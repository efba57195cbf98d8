// Block triangular solves  op(L) X = B  of the Nystrom preconditioner (M x M factor, M x dp right-hand
// sides, dp <= 32): HBM-bound -- one pass over the factor per solve -- with a sequential chain of
// M / 128 block steps on top.
//
// Replaces falkon `FalkonPreconditioner.invT / invTt / invA / invAt` (trsm) [external dependency],
// reached from /root/reference/sobolev_alignment/krr_approx.py:227 (fit) -- 4 solves per CG iteration.
//
// The factors are constant over the ~90 solves of a fit, so they are PACKED once (sa_trsm_pack):
//   * every strictly-lower 128x128 tile as an fp16 pair  x 2^s = hi + lo  (4 bytes per entry, the
//     traffic of fp32; ~22 bits), 64 KB per tile in exactly the shared-memory image the solve wants:
//     row r = [hi(r, 0..127) | lo(r, 0..127)] in 16-byte chunks, chunk index XOR (r & 7), so that one
//     `cp.async.bulk` per half tile lands it and every `ldmatrix` (plain for L x, `.trans` for L^T x)
//     is bank-conflict free.  Tile (r, c) sits at index r (r - 1) / 2 + c: the tiles of a block row
//     are contiguous and the packed factor is HALF the size of the fp32 square it replaces;
//   * per block the inverted diagonal block and the pre-multiplied sub-diagonal tiles
//       Wf[c] = inv(L_cc) L(c, c-1)     Wb[c] = L(c+1, c) inv(L_cc)
//     in the same format, so the block that was just published -- the last dependency of the next
//     one -- costs ONE tile product on the chain:
//       x_c = inv(L_cc) (b_c - sum_{k < c-1} L(c, k) x_k) - Wf[c] x_{c-1}.
//
// The solve is one persistent launch (one CTA per SM).  Work items = (job, block row, chunk of <= CH = 512
// dependency tiles); a CTA draws items from an atomic ticket in an order in which every item only waits
// for items with smaller tickets (chunk column by chunk column, rows top-down, the jobs of a pair
// interleaved), so heavy block rows are spread over many CTAs and the chain never waits for a CTA that
// is not resident.  Roles inside a CTA:
//   warp 8  (tile producer)    draws tickets, queues item descriptors, streams half tiles (32 KB
//                              `cp.async.bulk`, mbarrier complete_tx) through a 6-slot ring
//   warp 9  (x producer)       waits for the published-flag of each dependency block (ld.acquire.gpu)
//                              and bulk-copies its converted x (fp16 hi/lo, [column][k], swizzled, with
//                              its power-of-two scale in the trailer) into a 2-slot ring
//   warps 0-7 (consumers)      16 output rows each: ldmatrix + mma.sync.m16n8k16 (hi hi + lo hi + hi lo),
//                              fp32 accumulators in registers; the publisher of a block converts its x
//                              ONCE for all later consumers.
// Partial sums of a block row travel between its items through global memory in item order, and inside an
// item the dependency products are summed in groups of 16 (two-level summation: the rounding error does not
// grow with the length of a row): fixed summation order, so the result is bit-reproducible, which the
// row-sharded solver relies on.
//
// The chain itself -- block c needs x_{c-1} -- runs through a low-latency buffer: the publisher's threads write
// their fragment as 8-byte {payload, tag} words (relaxed, gpu scope) and the next rows poll exactly those words,
// so a chain step carries no fence, no flag round trip and no copy (see ll_store / ll_load below; measured anatomy
// of a step in profiles/r02_trsm_chain_trace.txt).  Older dependencies and all non-finishing items use the flag +
// bulk-copy route of warp 9.
//
// The same tile image and tile product serve `sa_ktv` (K^T t from a stored kernel block, the optional design (B) of
// the CG product pair): see the section in front of the packing kernels.
//
// sa_trsm_solve_pair runs the two dependent solves the FALKON operator always applies back to back
// (A^-1 then T^-1; T^-T, + lambda v, then A^-T) as two interleaved chains of one launch: job 1's block c
// takes job 0's published block c (+ lambda V) as its right-hand side.
#include <cuda_fp16.h>

#include <cstdlib>

#include "../../include/sobolev_b200.h"
#include "common.cuh"
#include "ptx.cuh"

namespace sab {
namespace tr {

constexpr int TS = 128;
constexpr int TILE_BYTES = TS * TS * 4;      // 64 KB: hi + lo
constexpr int HALF_BYTES = TILE_BYTES / 2;   // rows 0-63 / 64-127 of the stored tile
constexpr int ROW_BYTES = 512;
constexpr int NSLOT = 6;                     // half-tile ring: the last three tiles of a finishing row all fit
constexpr int NXSLOT = 2;                    // converted-x ring (one slot per tile; also the inverse product's operand)
constexpr int IQ = 4;                        // item queue depth
constexpr int NCONS = 8;                     // consumer warps
constexpr int THREADS = (NCONS + 2) * 32;
constexpr int SYNC_HDR = 64;                 // ints in front of the flags
enum { KIND_DEP = 0, KIND_INV = 1, KIND_W = 2 };
constexpr int LL_DEPS = 3;                   // a finishing row takes its last LL_DEPS dependencies (and the W tile's x) through
                                             // the low-latency buffer: see "publication" below
constexpr unsigned LL_TAG = 0x5AB00001u;

__host__ __device__ constexpr int xhat_bytes(int nb8) { return 2 * nb8 * 8 * 256 + 16; }
// 8-byte words {fp16 hi | fp16 lo, tag} of one converted block in the low-latency buffer (+ one word for its scale)
__host__ __device__ constexpr int ll_words(int nb8) { return TS * nb8 * 8 + 2; }

struct Job {
  const uint8_t* tiles;  // packed strictly-lower tiles
  const uint8_t* aux;    // [nblk][3] packed {inverse diagonal block, Wf, Wb}
  const float* scales;   // [3] 1 / operand scale of {tiles, inverse blocks, W tiles}
  const float* rhs;      // M x dp; job 0: == x (in place); job 1: job 0's x
  float* x;
  const float* V;        // optional additive term of the rhs (NULL = none)
  float lambda;
};
struct Params {
  Job job[2];
  int njobs, nblk, dp, chunk, nitems;
  int* sync;        // [0] ticket | per job j, [SYNC_HDR + (3 j + i) nblk + ..]: i = 0 converted x published (by block),
                    // 1 partial-sum count (by chain row), 2 fp32 x written (by block)
  float* partial;   // [njobs][nblk][128 x 32]
  uint8_t* xhat;    // [njobs][nblk][xhat_bytes]
  uint2* ll;        // [njobs][nblk][ll_words]: low-latency copy of the converted x (zeroed by the launch)
  long long* trace; // developer tool (option SAB_TRSM_TRACE_PTR): 8 globaltimer stamps per block of job 0
};

struct Item {
  int job, c, k0, ndep, q, fin;   // chain index c (forward: block c, backward: nblk-1-c); deps k0 .. k0+ndep-1
};

// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// bounded spin: a scheduling bug must surface as a trapped kernel, never as a hung GPU
__device__ __forceinline__ void wait_flag_ge(const int* f, int want, int what, unsigned backoff_ns) {
  if (ld_acquire_gpu(f) >= want) return;
  unsigned spins = 0;
  unsigned long long t0 = 0;
  while (ld_acquire_gpu(f) < want) {
    if (backoff_ns) __nanosleep(backoff_ns);
    if ((++spins & 0x3FFu) == 0) {
      const unsigned long long now = global_timer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ull) {
        printf("sobolev_b200: trsm wait timed out (cta %d thread %d waiting for %d >= %d)\n", blockIdx.x,
               threadIdx.x, what, want);
        __trap();
      }
    }
  }
}

// Low-latency publication (the idea of NCCL's LL protocol): every 8-byte word carries 4 bytes of payload and a tag,
// written with one relaxed (gpu scope) 8-byte store and polled with relaxed 8-byte loads -- single-copy atomic at that
// size, free to overlap each other (volatile accesses of a thread serialise: 12 words took 1.6 us to store and 4.4 us to
// read in the first version of this path) -- so the data is its own flag and the chain
// needs neither a fence on the producer side nor a separate flag round trip + copy on the consumer side.
__device__ __forceinline__ void ll_store(uint2* p, uint32_t payload) {
  asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(payload), "r"(LL_TAG) : "memory");
}
__device__ __forceinline__ uint2 ll_load(const uint2* p) {
  uint2 v;
  asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2(uint32_t addr, uint32_t (&r)[2]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// ticket -> item.  Order: chunk column j = 0, 1, ...; inside a column the rows that still have
// dependencies in it, top-down, the jobs interleaved row by row.  Column 0 also holds row 0.
__device__ __forceinline__ Item decode_item(int t, int njobs, int nblk, int CH) {
  int j = 0;
  long long base = 0;
  for (;;) {
    const int rows = j == 0 ? nblk : nblk - 1 - j * CH;
    const long long cnt = static_cast<long long>(rows) * njobs;
    if (t < base + cnt) break;
    base += cnt;
    ++j;
  }
  const int u = t - static_cast<int>(base);
  Item it;
  it.job = u % njobs;
  const int r = u / njobs;
  it.c = j == 0 ? r : j * CH + 1 + r;
  it.q = j;
  it.k0 = j * CH;
  it.fin = it.c <= (j + 1) * CH;
  const int k1 = it.fin ? it.c - 1 : (j + 1) * CH;   // the last dependency of a finishing row goes through W
  it.ndep = k1 > it.k0 ? k1 - it.k0 : 0;
  return it;
}
__host__ __device__ inline long long count_items(int njobs, int nblk, int CH) {
  long long n = 0;
  for (int j = 0;; ++j) {
    const int rows = j == 0 ? nblk : nblk - 1 - j * CH;
    if (rows <= 0) break;
    n += static_cast<long long>(rows) * njobs;
  }
  return n;
}

struct Smem {
  // barriers
  unsigned long long full[NSLOT], empty[NSLOT], xfull[NXSLOT], xempty[NXSLOT], iqfull[IQ], iqempty[IQ];
  Item iq[IQ];
  float red[2][NCONS];
  float tscale[2][4];   // 1 / operand scale of {tiles, inverse blocks, W tiles} per job
};

// 128x128 (half: 64 stored rows) tile product on the legacy tensor path: acc += hi x_hi, acx += lo x_hi + hi x_lo.
// `tile` = shared address of the half-tile slot, `lrow0` = first stored row (within the slot) of this k16 step /
// row block, xop = shared address of the converted operand ([n][k] hi, then lo at +NP*256).
template <int NB8, bool TRANS, int NSTEPS>
__device__ __forceinline__ void mma_steps(uint32_t tile, uint32_t xop, int kstep0, int m0_or_lrow,
                                          float (&acc)[NB8][4], float (&acx)[NB8][4]) {
  constexpr int NP = NB8 * 8;
  const int lane = threadIdx.x & 31;
  const int j = lane >> 3, i = lane & 7;
#pragma unroll
  for (int s = 0; s < NSTEPS; ++s) {
    const int ks = kstep0 + s;   // k16 step index within the 128-long contraction
    uint32_t ahi[4], alo[4];
    if (!TRANS) {
      // stored rows = output rows; this warp's 16 rows start at local row m0_or_lrow
      const int row = m0_or_lrow + (j & 1) * 8 + i;
      const int cc = ks * 2 + (j >> 1);
      const uint32_t rb = tile + row * ROW_BYTES;
      ldsm_x4(rb + ((cc ^ i) << 4), ahi);
      ldsm_x4(rb + (((16 + cc) ^ i) << 4), alo);
    } else {
      // stored rows = k; local stored row of this step inside the half slot, output rows m0 = m0_or_lrow
      const int row = (ks & 3) * 16 + (j >> 1) * 8 + i;
      const int cc = (m0_or_lrow >> 3) + (j & 1);
      const uint32_t rb = tile + row * ROW_BYTES;
      ldsm_x4_t(rb + ((cc ^ i) << 4), ahi);
      ldsm_x4_t(rb + (((16 + cc) ^ i) << 4), alo);
    }
#pragma unroll
    for (int nb = 0; nb + 1 < NB8; nb += 2) {
      const int n = (nb + (j >> 1)) * 8 + i;
      const int ccb = ks * 2 + (j & 1);
      const uint32_t a = xop + n * 256 + ((ccb ^ i) << 4);
      uint32_t bh[4], bl[4];
      ldsm_x4(a, bh);
      ldsm_x4(a + NP * 256, bl);
      mma_f16(acc[nb], ahi, bh[0], bh[1]);
      mma_f16(acx[nb], alo, bh[0], bh[1]);
      mma_f16(acx[nb], ahi, bl[0], bl[1]);
      mma_f16(acc[nb + 1], ahi, bh[2], bh[3]);
      mma_f16(acx[nb + 1], alo, bh[2], bh[3]);
      mma_f16(acx[nb + 1], ahi, bl[2], bl[3]);
    }
    if (NB8 & 1) {
      constexpr int nb = NB8 - 1;
      const int n = nb * 8 + i;
      const int ccb = ks * 2 + (j & 1);   // lanes 0-15 supply the addresses of .x2 (j = 0, 1)
      const uint32_t a = xop + n * 256 + ((ccb ^ i) << 4);
      uint32_t bh[2], bl[2];
      ldsm_x2(a, bh);
      ldsm_x2(a + NP * 256, bl);
      mma_f16(acc[nb], ahi, bh[0], bh[1]);
      mma_f16(acx[nb], alo, bh[0], bh[1]);
      mma_f16(acx[nb], ahi, bl[0], bl[1]);
    }
  }
}

// byte offset of half (n, k) inside one part (hi or lo) of a converted operand
__device__ __forceinline__ int xoff(int n, int k) { return n * 256 + ((((k >> 3) ^ (n & 7))) << 4) + (k & 7) * 2; }

template <int NB8, bool TRANS>
__global__ void __launch_bounds__(THREADS, 1) trsm_solve_kernel(const Params p) {
  constexpr int NP = NB8 * 8;
  constexpr int XB = xhat_bytes(NB8);
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* ring = smem_raw;                                    // NSLOT x 32 KB
  uint8_t* xring = ring + NSLOT * HALF_BYTES;                  // NXSLOT x XB (XB is a multiple of 16)
  Smem& sm = *reinterpret_cast<Smem*>(xring + NXSLOT * XB);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nblk = p.nblk, CH = p.chunk, dp = p.dp;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSLOT; ++s) {
      mbar_init(smem_u32(&sm.full[s]), 1);
      mbar_init(smem_u32(&sm.empty[s]), TRANS ? NCONS : NCONS / 2);
    }
    for (int s = 0; s < NXSLOT; ++s) {
      mbar_init(smem_u32(&sm.xfull[s]), 1);
      mbar_init(smem_u32(&sm.xempty[s]), NCONS);
    }
    for (int s = 0; s < IQ; ++s) {
      mbar_init(smem_u32(&sm.iqfull[s]), 1);
      mbar_init(smem_u32(&sm.iqempty[s]), NCONS + 1);
    }
    fence_barrier_init();
  }
  if (threadIdx.x < 8)
    sm.tscale[threadIdx.x >> 2][threadIdx.x & 3] =
        (static_cast<int>(threadIdx.x >> 2) < p.njobs && (threadIdx.x & 3) < 3)
            ? __ldg(p.job[threadIdx.x >> 2].scales + (threadIdx.x & 3)) : 0.f;
  __syncthreads();

  auto blk_of = [&](int chain) { return TRANS ? nblk - 1 - chain : chain; };
  // packed tile holding L(row blk of chain index c, dependency chain index k)
  auto dep_tile = [&](const Job& J, int c, int k) -> const uint8_t* {
    const long long r = TRANS ? nblk - 1 - k : c, cc = TRANS ? nblk - 1 - c : k;   // tile (r, cc), r > cc
    return J.tiles + (r * (r - 1) / 2 + cc) * static_cast<long long>(TILE_BYTES);
  };
  auto aux_tile = [&](const Job& J, int c, int kind) -> const uint8_t* {
    const int which = kind == KIND_INV ? 0 : (TRANS ? 2 : 1);
    return J.aux + (static_cast<long long>(blk_of(c)) * 3 + which) * TILE_BYTES;
  };
  auto ntiles_of = [&](const Item& it) { return it.ndep + (it.fin ? (it.c >= 1 ? 2 : 1) : 0); };
  auto kind_of = [&](const Item& it, int i) { return i < it.ndep ? KIND_DEP : (i == it.ndep ? KIND_INV : KIND_W); };

  if (warp == NCONS) {
    // ================================ tile producer ================================
    int g = 0;    // half-tile stage counter
    int nq = 0;   // items queued
    for (;;) {
      int t = 0;
      if (lane == 0) t = atomicAdd(p.sync, 1);
      t = __shfl_sync(0xffffffffu, t, 0);
      Item it;
      if (t < p.nitems) it = decode_item(t, p.njobs, nblk, CH);
      else { it.job = 0; it.c = -1; it.k0 = 0; it.ndep = 0; it.q = 0; it.fin = 0; }
      const int qs = nq % IQ;
      mbar_wait(smem_u32(&sm.iqempty[qs]), ((nq / IQ) & 1) ^ 1);
      if (lane == 0) {
        sm.iq[qs] = it;
        mbar_arrive(smem_u32(&sm.iqfull[qs]));
      }
      ++nq;
      if (it.c < 0) break;
      const Job& J = p.job[it.job];
      const int nt = ntiles_of(it);
      for (int i = 0; i < nt; ++i) {
        const int kd = kind_of(it, i);
        const uint8_t* src = kd == KIND_DEP ? dep_tile(J, it.c, it.k0 + i) : aux_tile(J, it.c, kd);
#pragma unroll
        for (int h = 0; h < 2; ++h, ++g) {
          const int s = g % NSLOT;
          mbar_wait(smem_u32(&sm.empty[s]), ((g / NSLOT) & 1) ^ 1);
          if (lane == 0) {
            mbar_arrive_expect_tx(smem_u32(&sm.full[s]), HALF_BYTES);
            bulk_load_1d(smem_u32(ring + s * HALF_BYTES), src + h * HALF_BYTES, HALF_BYTES, smem_u32(&sm.full[s]));
          }
        }
      }
    }
  } else if (warp == NCONS + 1) {
    // ================================ x producer ================================
    // One slot of the x ring per tile, in tile order.  Dependency / W tiles: wait for the block's published
    // flag, bulk-copy its converted x.  Inverse tile: the slot is handed to the consumers empty (they build
    // the operand b_c - sum themselves).
    int xg = 0, nq = 0;
    for (;;) {
      const int qs = nq % IQ;
      mbar_wait(smem_u32(&sm.iqfull[qs]), (nq / IQ) & 1);
      const Item it = sm.iq[qs];
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&sm.iqempty[qs]));
      ++nq;
      if (it.c < 0) break;
      const int nt = ntiles_of(it);
      const int* pub = p.sync + SYNC_HDR + (3 * it.job) * nblk;
      const uint8_t* xh = p.xhat + static_cast<long long>(it.job) * nblk * XB;
      for (int i = 0; i < nt; ++i, ++xg) {
        const int kd = kind_of(it, i);
        const int s = xg % NXSLOT;
        mbar_wait(smem_u32(&sm.xempty[s]), ((xg / NXSLOT) & 1) ^ 1);
        // tiles whose operand the consumers fetch themselves: the inverse product, and -- on a finishing row -- the
        // freshest dependencies, which come through the low-latency buffer
        const bool ll = it.fin && (kd == KIND_W || (kd == KIND_DEP && i >= it.ndep - LL_DEPS));
        if (lane == 0) {
          if (kd == KIND_INV || ll) {
            mbar_arrive(smem_u32(&sm.xfull[s]));
          } else {
            const int dep = kd == KIND_DEP ? it.k0 + i : it.c - 1;   // chain index of the block whose x is needed
            const int b = blk_of(dep);
            // only the W tile of a finishing row sits on the chain: everything else polls politely
            wait_flag_ge(pub + b, 1, b, kd == KIND_W ? 0u : 400u);
            fence_proxy_async_all();   // the block was written through the generic proxy of another SM
            if (p.trace != nullptr && kd == KIND_W && it.job == 0)
              p.trace[static_cast<long long>(blk_of(it.c)) * 8 + 0] = static_cast<long long>(global_timer_ns());
            mbar_arrive_expect_tx(smem_u32(&sm.xfull[s]), XB);
            bulk_load_1d(smem_u32(xring + s * XB), xh + static_cast<long long>(b) * XB, XB, smem_u32(&sm.xfull[s]));
          }
        }
        __syncwarp();
      }
    }
  } else {
    // ================================ consumers ================================
    const int g8 = lane >> 2, t4 = lane & 3;
    const int m0 = warp * 16;                  // output rows of this warp inside the block
    const int myhalf = warp >> 2;              // forward: which stored half holds this warp's rows
    int T = 0;                                 // tiles consumed (stage counter = 2 T + half; x slot = T % NXSLOT)
    int nq = 0, nred = 0;
    const uint32_t ring_a = smem_u32(ring), xring_a = smem_u32(xring);

    // block-wide max of |v| over the 128 x NP operand held in fragment layout
    auto block_max = [&](float m) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      float* red = sm.red[nred & 1];
      ++nred;
      if (lane == 0) red[warp] = m;
      consumer_bar();
      float r = red[0];
#pragma unroll
      for (int w = 1; w < NCONS; ++w) r = fmaxf(r, red[w]);
      return r;
    };
    // scale = 2^(14 - e) brings the largest entry into [2^13, 2^14); returns 1 / scale
    auto pick_scale = [&](float m, float& sc) {
      int e = 0;
      if (m > 0.f && m < 3.0e38f) frexpf(m, &e); else e = 14;
      sc = ldexpf(1.f, 14 - e);
      return ldexpf(1.f, e - 14);
    };
    // fragment (fp32, [128 x NP]) -> converted operand (fp16 hi / lo, [n][k] swizzled) in shared memory
    auto store_operand = [&](uint8_t* dst, const float (&v)[NB8][4], float sc) {
#pragma unroll
      for (int nb = 0; nb < NB8; ++nb)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int n = nb * 8 + 2 * t4 + (q & 1), k = m0 + g8 + (q >> 1) * 8;
          const float sv = v[nb][q] * sc;
          const __half h = __float2half_rn(sv);
          *reinterpret_cast<__half*>(dst + xoff(n, k)) = h;
          *reinterpret_cast<__half*>(dst + NP * 256 + xoff(n, k)) = __float2half_rn(sv - __half2float(h));
        }
    };
    auto stamp = [&](int blk, int i) {
      if (p.trace != nullptr && threadIdx.x == 0) p.trace[static_cast<long long>(blk) * 8 + i] = static_cast<long long>(global_timer_ns());
    };

    for (;;) {
      const int qs = nq % IQ;
      mbar_wait(smem_u32(&sm.iqfull[qs]), (nq / IQ) & 1);
      const Item it = sm.iq[qs];
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&sm.iqempty[qs]));
      ++nq;
      if (it.c < 0) break;
      const Job& J = p.job[it.job];
      const int blk = blk_of(it.c);
      const int nt = ntiles_of(it);
      const bool tr = p.trace != nullptr && it.fin && it.job == 0;
      int* pub = p.sync + SYNC_HDR + (3 * it.job) * nblk;
      int* pcnt = p.sync + SYNC_HDR + (3 * it.job + 1) * nblk;
      int* pub32 = p.sync + SYNC_HDR + (3 * it.job + 2) * nblk;
      float* part = p.partial + (static_cast<long long>(it.job) * nblk + it.c) * (TS * 32);
      const long long frow = static_cast<long long>(blk) * TS + m0 + g8;   // global row of fragment rows (+8)
      if (tr) stamp(blk, 7);

      // S: running sum over the eliminated dependencies; Sb: the current group of 16 -- two-level summation, so the
      // rounding error grows with 16 + (tiles / 16) additions instead of with the (up to 512) tiles of an item
      float S[NB8][4], Sb[NB8][4], U[NB8][4];
#pragma unroll
      for (int nb = 0; nb < NB8; ++nb)
#pragma unroll
        for (int q = 0; q < 4; ++q) S[nb][q] = Sb[nb][q] = U[nb][q] = 0.f;

      // S += the partial sum the previous item of this block row left behind (fixed order: item q-1 before q)
      auto add_previous_partial = [&]() {
        if (it.q == 0) return;
        wait_flag_ge(pcnt + it.c, it.q, -it.c, 0u);
#pragma unroll
        for (int nb = 0; nb < NB8; ++nb) {
          const int cc = nb * 8 + 2 * t4;
          const float2 v0 = __ldcg(reinterpret_cast<const float2*>(part + (m0 + g8) * 32 + cc));
          const float2 v1 = __ldcg(reinterpret_cast<const float2*>(part + (m0 + g8 + 8) * 32 + cc));
          S[nb][0] += v0.x; S[nb][1] += v0.y; S[nb][2] += v1.x; S[nb][3] += v1.y;
        }
      };
      // right-hand side fragment b_c (+ lambda V): job 0 reads its own input, job 1 the block job 0 published
      float bf[NB8][4];
      auto load_rhs = [&]() {
#pragma unroll
        for (int nb = 0; nb < NB8; ++nb) {
          const int cc = nb * 8 + 2 * t4;
          float2 b0 = make_float2(0.f, 0.f), b1 = make_float2(0.f, 0.f);
          if (cc < dp) {
            b0 = __ldcg(reinterpret_cast<const float2*>(J.rhs + frow * dp + cc));
            b1 = __ldcg(reinterpret_cast<const float2*>(J.rhs + (frow + 8) * dp + cc));
            if (J.V != nullptr) {
              const float2 a0 = *reinterpret_cast<const float2*>(J.V + frow * dp + cc);
              const float2 a1 = *reinterpret_cast<const float2*>(J.V + (frow + 8) * dp + cc);
              b0.x = fmaf(J.lambda, a0.x, b0.x); b0.y = fmaf(J.lambda, a0.y, b0.y);
              b1.x = fmaf(J.lambda, a1.x, b1.x); b1.y = fmaf(J.lambda, a1.y, b1.y);
            }
          }
          bf[nb][0] = b0.x; bf[nb][1] = b0.y; bf[nb][2] = b1.x; bf[nb][3] = b1.y;
        }
      };
      // a finishing row fetches what does not depend on the chain up front: its earlier partial sum (written
      // a whole chunk column ago) and, for job 0, its right-hand side
      if (it.fin) {
        add_previous_partial();
        if (it.job == 0) load_rhs();
      }

      for (int i = 0; i < nt; ++i, ++T) {
        const int kd = kind_of(it, i);
        const int xs = T % NXSLOT;
        uint8_t* xslot = xring + xs * XB;
        const uint32_t xop = xring_a + xs * XB;
        float xinv;
        mbar_wait(smem_u32(&sm.xfull[xs]), (T / NXSLOT) & 1);
        if (kd == KIND_INV) {
          // operand = b_c - (sum of all eliminated dependencies), converted by this CTA into the handed-over slot
          if (it.job == 1) {
            wait_flag_ge(p.sync + SYNC_HDR + 2 * nblk + blk, 1, blk, 0u);   // job 0 has written this block (fp32)
            load_rhs();
          }
          float v[NB8][4];
          float m = 0.f;
#pragma unroll
          for (int nb = 0; nb < NB8; ++nb)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              v[nb][q] = bf[nb][q] - S[nb][q];
              m = fmaxf(m, fabsf(v[nb][q]));
            }
          float sc;
          xinv = pick_scale(block_max(m), sc);
          store_operand(xslot, v, sc);
          fence_proxy_async_smem();   // the slot is refilled by cp.async.bulk later: order these generic writes before it
          consumer_bar();
        } else if (it.fin && (kd == KIND_W || i >= it.ndep - LL_DEPS)) {
          // low-latency fetch: every thread polls the 4 NB8 words of ITS fragment positions (same thread -> word map
          // as the publisher) plus the scale word, and drops the halves into the handed-over slot
          const int dep = kd == KIND_DEP ? it.k0 + i : it.c - 1;
          const uint2* src = p.ll + (static_cast<long long>(it.job) * nblk + blk_of(dep)) * ll_words(NB8);
          // word (thread, e) sits at e * 256 + thread: a warp's access is 256 contiguous bytes (element-major; the
          // thread-major layout of the first version cost 32 partial sectors per instruction: 4.3 us per read)
          const uint2* mine = src + threadIdx.x;
          // every round re-issues ALL loads of the thread at once (one L2 round trip per round, not one per stale word)
          uint2 w[4 * NB8], ws;
          {
            unsigned spins = 0;
            unsigned long long t0 = 0;
            for (;;) {
#pragma unroll
              for (int e = 0; e < 4 * NB8; ++e) w[e] = ll_load(mine + e * (NCONS * 32));
              ws = ll_load(src + TS * NP);
              bool ok = ws.y == LL_TAG;
#pragma unroll
              for (int e = 0; e < 4 * NB8; ++e) ok = ok && (w[e].y == LL_TAG);
              if (ok) break;
              if ((++spins & 0x3FFu) == 0) {
                const unsigned long long now = global_timer_ns();
                if (t0 == 0) t0 = now;
                else if (now - t0 > 4000000000ull) {
                  printf("sobolev_b200: trsm low-latency wait timed out (cta %d thread %d block %d)\n", blockIdx.x,
                         threadIdx.x, blk_of(dep));
                  __trap();
                }
              }
            }
          }
#pragma unroll
          for (int e = 0; e < 4 * NB8; ++e) {
            const uint32_t pay = w[e].x;
            const int nb = e >> 2, q = e & 3;
            const int n = nb * 8 + 2 * t4 + (q & 1), k = m0 + g8 + (q >> 1) * 8;
            *reinterpret_cast<uint16_t*>(xslot + xoff(n, k)) = static_cast<uint16_t>(pay & 0xffffu);
            *reinterpret_cast<uint16_t*>(xslot + NP * 256 + xoff(n, k)) = static_cast<uint16_t>(pay >> 16);
          }
          xinv = __uint_as_float(ws.x);
          if (tr && kd == KIND_W) stamp(blk, 1);
          fence_proxy_async_smem();
          consumer_bar();
        } else {
          xinv = *reinterpret_cast<const float*>(xslot + 2 * NP * 256);
        }
        float acc[NB8][4], acx[NB8][4];
#pragma unroll
        for (int nb = 0; nb < NB8; ++nb)
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[nb][q] = acx[nb][q] = 0.f;
        if (!TRANS) {
          const int g = 2 * T + myhalf, s = g % NSLOT;
          mbar_wait(smem_u32(&sm.full[s]), (g / NSLOT) & 1);
          if (tr && kd == KIND_W) stamp(blk, 2);
          mma_steps<NB8, false, 8>(ring_a + s * HALF_BYTES, xop, 0, m0 - myhalf * 64, acc, acx);
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&sm.empty[s]));
        } else {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int g = 2 * T + h, s = g % NSLOT;
            mbar_wait(smem_u32(&sm.full[s]), (g / NSLOT) & 1);
            if (tr && kd == KIND_W && h == 1) stamp(blk, 2);
            mma_steps<NB8, true, 4>(ring_a + s * HALF_BYTES, xop, h * 4, m0, acc, acx);
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&sm.empty[s]));
          }
        }
        // the x slot of a finishing row's LAST tile is kept: it stages the converted x_c for publication
        if (!(it.fin && i == nt - 1)) {
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&sm.xempty[xs]));
        }
        const float r = sm.tscale[it.job][kd] * xinv;
#pragma unroll
        for (int nb = 0; nb < NB8; ++nb)
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float pv = (acc[nb][q] + acx[nb][q]) * r;
            if (kd == KIND_DEP) Sb[nb][q] += pv;
            else if (kd == KIND_INV) U[nb][q] = pv;
            else U[nb][q] -= pv;
          }
        if (kd == KIND_DEP && (((it.k0 + i) & 15) == 15 || i + 1 == it.ndep)) {
#pragma unroll
          for (int nb = 0; nb < NB8; ++nb)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              S[nb][q] += Sb[nb][q];
              Sb[nb][q] = 0.f;
            }
        }
      }

      if (!it.fin) {
        // hand the running sum of this block row to its next item
        add_previous_partial();
#pragma unroll
        for (int nb = 0; nb < NB8; ++nb) {
          const int cc = nb * 8 + 2 * t4;
          __stcg(reinterpret_cast<float2*>(part + (m0 + g8) * 32 + cc), make_float2(S[nb][0], S[nb][1]));
          __stcg(reinterpret_cast<float2*>(part + (m0 + g8 + 8) * 32 + cc), make_float2(S[nb][2], S[nb][3]));
        }
        consumer_bar();
        if (threadIdx.x == 0) {
          __threadfence();
          st_release_gpu(pcnt + it.c, it.q + 1);
        }
      } else {
        // x_c: the converted operand every later block row waits for goes out FIRST (it is the chain); the
        // fp32 result (final output, and job 1's right-hand side) follows behind a flag of its own
        if (tr) stamp(blk, 3);
        float m = 0.f;
#pragma unroll
        for (int nb = 0; nb < NB8; ++nb)
#pragma unroll
          for (int q = 0; q < 4; ++q) m = fmaxf(m, fabsf(U[nb][q]));
        float sc;
        const float inv = pick_scale(block_max(m), sc);   // (its barrier: every warp is done reading the slot)
        if (tr) stamp(blk, 4);
        {
          // the chain: 4 NB8 tagged words per thread straight from the registers, no barrier, no fence
          uint2* dst = p.ll + (static_cast<long long>(it.job) * nblk + blk) * ll_words(NB8);
          uint2* mine = dst + threadIdx.x;
#pragma unroll
          for (int nb = 0; nb < NB8; ++nb)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float sv = U[nb][q] * sc;
              const __half h = __float2half_rn(sv);
              const __half l = __float2half_rn(sv - __half2float(h));
              ll_store(mine + (nb * 4 + q) * (NCONS * 32), static_cast<uint32_t>(__half_as_ushort(h)) |
                                              (static_cast<uint32_t>(__half_as_ushort(l)) << 16));
            }
          if (threadIdx.x == 0) ll_store(dst + TS * NP, __float_as_uint(inv));
          if (tr) stamp(blk, 6);
        }
        const int xs = (T - 1) % NXSLOT;
        uint8_t* stage = xring + xs * XB;
        store_operand(stage, U, sc);
        if (threadIdx.x == 0) *reinterpret_cast<float*>(stage + 2 * NP * 256) = inv;
        fence_proxy_async_smem();
        consumer_bar();
        // coalesced copy of the staged block to global memory (16 bytes per thread and step)
        uint8_t* xh = p.xhat + (static_cast<long long>(it.job) * nblk + blk) * XB;
        for (int o = threadIdx.x * 16; o < XB; o += NCONS * 32 * 16)
          __stcg(reinterpret_cast<uint4*>(xh + o), *reinterpret_cast<const uint4*>(stage + o));
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&sm.xempty[xs]));
        if (tr) stamp(blk, 5);
        consumer_bar();
        if (threadIdx.x == 0) {
          __threadfence();
          fence_proxy_async_all();
          st_release_gpu(pub + blk, 1);
        }
#pragma unroll
        for (int nb = 0; nb < NB8; ++nb) {
          const int cc = nb * 8 + 2 * t4;
          if (cc < dp) {
            __stcg(reinterpret_cast<float2*>(J.x + frow * dp + cc), make_float2(U[nb][0], U[nb][1]));
            __stcg(reinterpret_cast<float2*>(J.x + (frow + 8) * dp + cc), make_float2(U[nb][2], U[nb][3]));
          }
        }
        if (p.njobs > 1 && it.job == 0) {
          consumer_bar();
          if (threadIdx.x == 0) {
            __threadfence();
            st_release_gpu(pub32 + blk, 1);
          }
        }
      }
    }
  }
}

// ----------------------------------------------------------------------------------------------
// K^T t from a stored kernel block (the second half of a CG iteration's K_nM^T (K_nM z))
// ----------------------------------------------------------------------------------------------
// sa_kmv_store leaves k(X_blk, C) behind as packed 128 x 128 hi/lo tiles (the image above; entries carry
// 2^14, the residual another 2^11).  out[j, :] += sum_i k(x_i, c_j) t[i, :] is then the transposed tile
// product of the backward solve without any dependency: HBM-bound streaming of the block (4 bytes per
// kernel entry, once) instead of forming K a second time on the tensor cores (2 p flops per entry).
// t is converted per 128-row block like a published x (fp16 hi / lo, lo x 2^11, power-of-two block scale).
struct KtvParams {
  const uint8_t* tiles;   // [ni][nj] packed tiles
  const uint8_t* that;    // [ni][xhat_bytes] converted t blocks
  float* out;             // [b x dp], accumulated into (red.add)
  long long b;
  int ni, nj, dp, chunk;  // chunk = row blocks per work item
};

template <int NB8>
__global__ void __launch_bounds__(256) that_kernel(const float* __restrict__ t, long long a, int dp,
                                                   uint8_t* __restrict__ that) {
  constexpr int NP = NB8 * 8;
  constexpr int XB = xhat_bytes(NB8);
  __shared__ float red[8];
  const long long r0 = static_cast<long long>(blockIdx.x) * TS;
  uint8_t* dst = that + static_cast<long long>(blockIdx.x) * XB;
  float m = 0.f;
  for (int idx = threadIdx.x; idx < TS * dp; idx += 256) {
    const long long r = r0 + idx / dp;
    if (r < a) m = fmaxf(m, fabsf(t[r * dp + idx % dp]));
  }
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  m = red[0];
  for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]);
  int e = 14;
  if (m > 0.f && m < 3.0e38f) frexpf(m, &e);
  const float sc = ldexpf(1.f, 14 - e);
  for (int idx = threadIdx.x; idx < TS * NP; idx += 256) {
    const int k = idx / NP, n = idx % NP;
    const long long r = r0 + k;
    const float v = (n < dp && r < a) ? t[r * dp + n] * sc : 0.f;
    const __half h = __float2half_rn(v);
    *reinterpret_cast<__half*>(dst + xoff(n, k)) = h;
    *reinterpret_cast<__half*>(dst + NP * 256 + xoff(n, k)) = __float2half_rn((v - __half2float(h)) * 2048.f);
  }
  if (threadIdx.x == 0) *reinterpret_cast<float*>(dst + 2 * NP * 256) = ldexpf(1.f, e - 14);
}

struct KtvSmem {
  unsigned long long full[NSLOT], empty[NSLOT], xfull[NXSLOT], xempty[NXSLOT];
};

template <int NB8>
__global__ void __launch_bounds__(THREADS, 1) ktv_kernel(const KtvParams p) {
  constexpr int NP = NB8 * 8;
  constexpr int XB = xhat_bytes(NB8);
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* ring = smem_raw;
  uint8_t* xring = ring + NSLOT * HALF_BYTES;
  KtvSmem& sm = *reinterpret_cast<KtvSmem*>(xring + NXSLOT * XB);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NSLOT; ++s) {
      mbar_init(smem_u32(&sm.full[s]), 1);
      mbar_init(smem_u32(&sm.empty[s]), NCONS);
    }
    for (int s = 0; s < NXSLOT; ++s) {
      mbar_init(smem_u32(&sm.xfull[s]), 1);
      mbar_init(smem_u32(&sm.xempty[s]), NCONS);
    }
    fence_barrier_init();
  }
  __syncthreads();
  // items: (row chunk ic, column block jb), jb fastest -- the CTAs running together read the same converted
  // t blocks (L2 resident) against different tile columns
  const int nchunks = (p.ni + p.chunk - 1) / p.chunk;
  const long long nitems = static_cast<long long>(nchunks) * p.nj;
  if (warp == NCONS) {
    int g = 0;
    for (long long id = blockIdx.x; id < nitems; id += gridDim.x) {
      const int ic = static_cast<int>(id / p.nj), jb = static_cast<int>(id % p.nj);
      const int i1 = min(p.ni, (ic + 1) * p.chunk);
      for (int ib = ic * p.chunk; ib < i1; ++ib) {
        const uint8_t* src = p.tiles + (static_cast<long long>(ib) * p.nj + jb) * TILE_BYTES;
#pragma unroll
        for (int h = 0; h < 2; ++h, ++g) {
          const int s = g % NSLOT;
          mbar_wait(smem_u32(&sm.empty[s]), ((g / NSLOT) & 1) ^ 1);
          if (lane == 0) {
            mbar_arrive_expect_tx(smem_u32(&sm.full[s]), HALF_BYTES);
            bulk_load_1d(smem_u32(ring + s * HALF_BYTES), src + h * HALF_BYTES, HALF_BYTES, smem_u32(&sm.full[s]));
          }
        }
      }
    }
  } else if (warp == NCONS + 1) {
    int xg = 0;
    for (long long id = blockIdx.x; id < nitems; id += gridDim.x) {
      const int ic = static_cast<int>(id / p.nj);
      const int i1 = min(p.ni, (ic + 1) * p.chunk);
      for (int ib = ic * p.chunk; ib < i1; ++ib, ++xg) {
        const int s = xg % NXSLOT;
        mbar_wait(smem_u32(&sm.xempty[s]), ((xg / NXSLOT) & 1) ^ 1);
        if (lane == 0) {
          mbar_arrive_expect_tx(smem_u32(&sm.xfull[s]), XB);
          bulk_load_1d(smem_u32(xring + s * XB), p.that + static_cast<long long>(ib) * XB, XB, smem_u32(&sm.xfull[s]));
        }
        __syncwarp();
      }
    }
  } else {
    const int g8 = lane >> 2, t4 = lane & 3;
    const int m0 = warp * 16;
    const uint32_t ring_a = smem_u32(ring), xring_a = smem_u32(xring);
    int T = 0;
    for (long long id = blockIdx.x; id < nitems; id += gridDim.x) {
      const int ic = static_cast<int>(id / p.nj), jb = static_cast<int>(id % p.nj);
      const int i1 = min(p.ni, (ic + 1) * p.chunk);
      float S[NB8][4];
#pragma unroll
      for (int nb = 0; nb < NB8; ++nb)
#pragma unroll
        for (int q = 0; q < 4; ++q) S[nb][q] = 0.f;
      for (int ib = ic * p.chunk; ib < i1; ++ib, ++T) {
        const int xs = T % NXSLOT;
        mbar_wait(smem_u32(&sm.xfull[xs]), (T / NXSLOT) & 1);
        const uint32_t xop = xring_a + xs * XB;
        const float xinv = *reinterpret_cast<const float*>(xring + xs * XB + 2 * NP * 256);
        float acc[NB8][4], acx[NB8][4];
#pragma unroll
        for (int nb = 0; nb < NB8; ++nb)
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[nb][q] = acx[nb][q] = 0.f;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int g = 2 * T + h, s = g % NSLOT;
          mbar_wait(smem_u32(&sm.full[s]), (g / NSLOT) & 1);
          mma_steps<NB8, true, 4>(ring_a + s * HALF_BYTES, xop, h * 4, m0, acc, acx);
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&sm.empty[s]));
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&sm.xempty[xs]));
#pragma unroll
        for (int nb = 0; nb < NB8; ++nb)
#pragma unroll
          for (int q = 0; q < 4; ++q) S[nb][q] = fmaf(fmaf(acx[nb][q], 1.f / 2048.f, acc[nb][q]), xinv, S[nb][q]);
      }
      // entries of the stored block carry 2^14
      const long long row = static_cast<long long>(jb) * TS + m0 + g8;
#pragma unroll
      for (int nb = 0; nb < NB8; ++nb) {
        const int cc = nb * 8 + 2 * t4;
        if (cc < p.dp) {
          if (row < p.b) {
            red_add_f32(p.out + row * p.dp + cc, S[nb][0] * (1.f / 16384.f));
            red_add_f32(p.out + row * p.dp + cc + 1, S[nb][1] * (1.f / 16384.f));
          }
          if (row + 8 < p.b) {
            red_add_f32(p.out + (row + 8) * p.dp + cc, S[nb][2] * (1.f / 16384.f));
            red_add_f32(p.out + (row + 8) * p.dp + cc + 1, S[nb][3] * (1.f / 16384.f));
          }
        }
      }
    }
  }
}

// ----------------------------------------------------------------------------------------------
// packing
// ----------------------------------------------------------------------------------------------
__global__ void absmax_lower_kernel(const float* __restrict__ A, long long M, long long lda, unsigned* out) {
  float m = 0.f;
  for (long long r = blockIdx.x; r < M; r += gridDim.x) {
    const long long cend = (r / TS) * TS;   // strictly-lower TILES only
    for (long long c = threadIdx.x; c < cend; c += blockDim.x) m = fmaxf(m, fabsf(A[r * lda + c]));
  }
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));
}
// tiles [count][128*128] fp32, tile t taken when (t % 3) is in `mask` (bit per kind)
__global__ void absmax_tiles_kernel(const float* __restrict__ A, long long count, int stride_tiles, int first,
                                    unsigned* out) {
  float m = 0.f;
  for (long long t = blockIdx.x; t < count; t += gridDim.x) {
    const float* src = A + (t * stride_tiles + first) * (TS * TS);
    for (int idx = threadIdx.x; idx < TS * TS; idx += blockDim.x) m = fmaxf(m, fabsf(src[idx]));
  }
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));
}
// bits[0..2] -> scales[0..2] = 2^-(13 - e): the INVERSE of the operand scale 2^(13 - e) that brings the
// largest entry of the class into [2^12, 2^13); scales[4..6] = the scale itself
__global__ void pack_scales_kernel(const unsigned* bits, float* scales) {
  const int i = threadIdx.x;
  if (i < 3) {
    const float m = __uint_as_float(bits[i]);
    int e = 0;
    if (m > 0.f && m < 3.0e38f) frexpf(m, &e);
    scales[i] = ldexpf(1.f, e - 13);
    scales[4 + i] = ldexpf(1.f, 13 - e);
  }
}

// one CTA per tile, one warp per stored row at a time: lane l converts columns 4l .. 4l+3
__device__ __forceinline__ void pack_tile(const float* __restrict__ src, long long ld, float sc, uint8_t* __restrict__ dst) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < TS; r += blockDim.x >> 5) {
    const float4 v = *reinterpret_cast<const float4*>(src + static_cast<long long>(r) * ld + lane * 4);
    const float f[4] = {v.x * sc, v.y * sc, v.z * sc, v.w * sc};
    __half h[4], l[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      h[q] = __float2half_rn(f[q]);
      l[q] = __float2half_rn(f[q] - __half2float(h[q]));
    }
    const int cc = lane >> 1, sub = (lane & 1) * 8;
    uint8_t* row = dst + r * ROW_BYTES;
    *reinterpret_cast<uint2*>(row + ((cc ^ (r & 7)) << 4) + sub) = *reinterpret_cast<const uint2*>(h);
    *reinterpret_cast<uint2*>(row + (((16 + cc) ^ (r & 7)) << 4) + sub) = *reinterpret_cast<const uint2*>(l);
  }
}
__global__ void __launch_bounds__(256) pack_lower_kernel(const float* __restrict__ L, long long ldl, int nblk,
                                                          const float* __restrict__ scales, uint8_t* __restrict__ tiles) {
  // blockIdx.x = linear index of the strictly-lower tile (r, c) = r (r - 1) / 2 + c
  const long long t = blockIdx.x;
  long long r = static_cast<long long>((1.0 + sqrt(1.0 + 8.0 * static_cast<double>(t))) * 0.5);
  while (r * (r - 1) / 2 > t) --r;
  while ((r + 1) * r / 2 <= t) ++r;
  const long long c = t - r * (r - 1) / 2;
  pack_tile(L + r * TS * ldl + c * TS, ldl, scales[4 + 0], tiles + t * TILE_BYTES);
}
__global__ void __launch_bounds__(256) pack_aux_kernel(const float* __restrict__ invdiag, const float* __restrict__ W,
                                                        const float* __restrict__ scales, uint8_t* __restrict__ aux) {
  const int blk = blockIdx.x, which = blockIdx.y;   // 0 inverse block, 1 Wf, 2 Wb
  const float* src = which == 0 ? invdiag + static_cast<long long>(blk) * TS * TS
                                : W + (static_cast<long long>(blk) * 2 + (which - 1)) * TS * TS;
  pack_tile(src, TS, scales[4 + (which == 0 ? 1 : 2)], aux + (static_cast<long long>(blk) * 3 + which) * TILE_BYTES);
}

// W[c][0] = inv(L_cc) L(c, c-1),  W[c][1] = L(c+1, c) inv(L_cc)   (fp32 SIMT, once per factor)
constexpr int PM_LDA = TS + 1;
__global__ void __launch_bounds__(256)
premul_kernel(const float* __restrict__ L, long long ldl, const float* __restrict__ invdiag, int nblk,
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

// dependency tiles per work item (tuning knob SAB_TRSM_CHUNK).  Measured on B200 (tools/gpu_sweep.sh, pair at
// M = 50 048 / 100 096): 16 -> 2.65 / 8.4 ms, 64 -> 2.4 / 7.5, 256 -> 2.1 / 6.9, 512 -> 2.1 / 6.7, 1024 -> 2.1 / 7.0:
// every chunk-column boundary stalls the chain for one bulk item, so long chunks win until the rows of the very
// largest factors stop balancing.
static int chunk_tiles() {
  const int c = option_int("SAB_TRSM_CHUNK", 512);
  return c < 2 ? 2 : (c > 4096 ? 4096 : c);
}
static int sm_count() {
  static const int v = [] {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
    return n;
  }();
  return v;
}

static long long tri_tiles(long long nblk) { return nblk * (nblk - 1) / 2; }

template <int NB8>
static int launch_solve(cudaStream_t st, const Params& p, int transpose) {
  const int smem = NSLOT * HALF_BYTES + NXSLOT * xhat_bytes(NB8) + static_cast<int>(sizeof(Smem));
  const long long grid = p.nitems < sm_count() ? p.nitems : sm_count();
  auto go = [&](auto kern) -> int {
    SAB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kern<<<static_cast<int>(grid), THREADS, smem, st>>>(p);
    SAB_CHECK_CUDA(cudaGetLastError());
    return 0;
  };
  return transpose ? go(trsm_solve_kernel<NB8, true>) : go(trsm_solve_kernel<NB8, false>);
}

}  // namespace tr
}  // namespace sab

using namespace sab;
using namespace sab::tr;

extern "C" int64_t sa_trsm_packed_bytes(int64_t M) {
  if (M <= 0 || M % TS != 0) return -1;
  const int64_t nblk = M / TS;
  return (tri_tiles(nblk) + 3 * nblk) * TILE_BYTES + 256;
}

extern "C" int64_t sa_trsm_pack_workspace_bytes(int64_t M) {
  if (M <= 0 || M % TS != 0) return -1;
  return (M / TS) * 2 * static_cast<int64_t>(TS) * TS * 4 + 256;
}

extern "C" int sa_trsm_pack(void* stream, const float* L, int64_t M, int64_t ldl, const float* invdiag,
                            void* packed, void* workspace, int64_t workspace_bytes) {
  SAB_REQUIRE(M > 0 && M % TS == 0 && ldl >= M && ldl % 4 == 0 && (reinterpret_cast<uintptr_t>(L) & 15) == 0,
              "matrix order must be a positive multiple of 128 with ld %% 4 == 0 and 16-byte aligned storage");
  SAB_REQUIRE(invdiag != nullptr && packed != nullptr && (reinterpret_cast<uintptr_t>(packed) & 127) == 0,
              "packed buffer must be 128-byte aligned");
  SAB_REQUIRE(workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0 &&
                  workspace_bytes >= sa_trsm_pack_workspace_bytes(M),
              "workspace must hold sa_trsm_pack_workspace_bytes(M) = %lld bytes",
              (long long)sa_trsm_pack_workspace_bytes(M));
  SAB_REQUIRE(M / TS <= 46000, "matrix order %lld too large", (long long)M);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nblk = static_cast<int>(M / TS);
  uint8_t* tiles = static_cast<uint8_t*>(packed);
  uint8_t* aux = tiles + tri_tiles(nblk) * TILE_BYTES;
  float* scales = reinterpret_cast<float*>(aux + 3ll * nblk * TILE_BYTES);
  float* W = static_cast<float*>(workspace);
  unsigned* bits = reinterpret_cast<unsigned*>(scales) + 8;
  SAB_CHECK_CUDA(cudaMemsetAsync(scales, 0, 256, st));
  const int psmem = (TS * PM_LDA + TS * TS) * static_cast<int>(sizeof(float));
  SAB_CHECK_CUDA(cudaFuncSetAttribute(premul_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, psmem));
  premul_kernel<<<dim3(nblk, 2), 256, psmem, st>>>(L, ldl, invdiag, nblk, W);
  absmax_lower_kernel<<<1184, 256, 0, st>>>(L, M, ldl, bits + 0);
  absmax_tiles_kernel<<<nblk < 1184 ? nblk : 1184, 256, 0, st>>>(invdiag, nblk, 1, 0, bits + 1);
  absmax_tiles_kernel<<<2 * nblk < 1184 ? 2 * nblk : 1184, 256, 0, st>>>(W, 2ll * nblk, 1, 0, bits + 2);
  pack_scales_kernel<<<1, 32, 0, st>>>(bits, scales);
  if (nblk > 1) pack_lower_kernel<<<static_cast<unsigned>(tri_tiles(nblk)), 256, 0, st>>>(L, ldl, nblk, scales, tiles);
  pack_aux_kernel<<<dim3(nblk, 3), 256, 0, st>>>(invdiag, W, scales, aux);
  SAB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int64_t sa_trsm_workspace_bytes(int64_t M) {
  if (M <= 0 || M % TS != 0) return -1;
  const int64_t nblk = M / TS;
  int64_t b = (SYNC_HDR + 6 * nblk) * 4;
  b = (b + 255) / 256 * 256;
  b += 2 * nblk * TS * 32 * 4;                 // partial sums
  b += 2 * nblk * ((xhat_bytes(4) + 127) / 128 * 128);   // converted x blocks
  b += 2 * nblk * static_cast<int64_t>(ll_words(4)) * 8;     // their low-latency copies
  return b + 256;
}

static int solve_jobs(cudaStream_t st, Params& p, int64_t M, int dp, int transpose, void* workspace,
                      int64_t workspace_bytes) {
  SAB_REQUIRE(workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & 127) == 0 &&
                  workspace_bytes >= sa_trsm_workspace_bytes(M),
              "workspace must be 128-byte aligned and hold sa_trsm_workspace_bytes(M) = %lld bytes",
              (long long)sa_trsm_workspace_bytes(M));
  SAB_REQUIRE(dp >= 4 && dp <= 32 && dp % 4 == 0, "dp must be a multiple of 4 in [4, 32], got %d", dp);
  const int nblk = static_cast<int>(M / TS);
  p.nblk = nblk;
  p.dp = dp;
  p.chunk = chunk_tiles();
  p.nitems = static_cast<int>(count_items(p.njobs, nblk, p.chunk));
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  p.sync = reinterpret_cast<int*>(ws);
  int64_t off = (SYNC_HDR + 6ll * nblk) * 4;
  off = (off + 255) / 256 * 256;
  p.partial = reinterpret_cast<float*>(ws + off);
  off += 2ll * nblk * TS * 32 * 4;
  p.xhat = ws + off;
  off += 2ll * nblk * ((xhat_bytes(4) + 127) / 128 * 128);
  p.ll = reinterpret_cast<uint2*>(ws + off);
  SAB_CHECK_CUDA(cudaMemsetAsync(p.ll, 0, static_cast<size_t>(p.njobs) * nblk * ll_words(4) * 8, st));
  {
    char tp[64];
    p.trace = option("SAB_TRSM_TRACE_PTR", tp) ? reinterpret_cast<long long*>(std::strtoull(tp, nullptr, 0)) : nullptr;
  }
  SAB_CHECK_CUDA(cudaMemsetAsync(p.sync, 0, (SYNC_HDR + 6ll * nblk) * 4, st));
  switch ((dp + 7) / 8) {
    case 1: return launch_solve<1>(st, p, transpose);
    case 2: return launch_solve<2>(st, p, transpose);
    case 3: return launch_solve<3>(st, p, transpose);
    default: return launch_solve<4>(st, p, transpose);
  }
}

static Job make_job(const void* packed, int64_t M, const float* rhs, float* x, const float* V, float lambda) {
  const int64_t nblk = M / TS;
  const uint8_t* tiles = static_cast<const uint8_t*>(packed);
  const uint8_t* aux = tiles + tri_tiles(nblk) * TILE_BYTES;
  Job j;
  j.tiles = tiles;
  j.aux = aux;
  j.scales = reinterpret_cast<const float*>(aux + 3 * nblk * TILE_BYTES);
  j.rhs = rhs;
  j.x = x;
  j.V = V;
  j.lambda = lambda;
  return j;
}

extern "C" int sa_trsm_solve(void* stream, const void* packed, int64_t M, float* B, int dp, int transpose,
                             void* workspace, int64_t workspace_bytes) {
  SAB_REQUIRE(M > 0 && M % TS == 0 && packed != nullptr && B != nullptr, "bad operands");
  Params p = {};
  p.njobs = 1;
  p.job[0] = make_job(packed, M, B, B, nullptr, 0.f);
  return solve_jobs(static_cast<cudaStream_t>(stream), p, M, dp, transpose, workspace, workspace_bytes);
}

extern "C" int sa_trsm_solve_pair(void* stream, const void* packed_a, const void* packed_b, int64_t M, float* Xa,
                                  float* Xb, const float* V, float lambda, int dp, int transpose, void* workspace,
                                  int64_t workspace_bytes) {
  SAB_REQUIRE(M > 0 && M % TS == 0 && packed_a != nullptr && packed_b != nullptr && Xa != nullptr &&
                  Xb != nullptr && Xa != Xb,
              "bad operands");
  Params p = {};
  p.njobs = 2;
  p.job[0] = make_job(packed_a, M, Xa, Xa, nullptr, 0.f);
  p.job[1] = make_job(packed_b, M, Xa, Xb, V, lambda);
  return solve_jobs(static_cast<cudaStream_t>(stream), p, M, dp, transpose, workspace, workspace_bytes);
}

// ---- K^T t from a stored kernel block
extern "C" int64_t sa_ktv_workspace_bytes(int64_t a) {
  if (a <= 0) return -1;
  return ((a + TS - 1) / TS) * static_cast<int64_t>((xhat_bytes(4) + 127) / 128 * 128) + 256;
}

template <int NB8>
static int ktv_launch(cudaStream_t st, const float* t, int64_t a, KtvParams p, uint8_t* that) {
  that_kernel<NB8><<<p.ni, 256, 0, st>>>(t, a, p.dp, that);
  SAB_CHECK_CUDA(cudaGetLastError());
  const int smem = NSLOT * HALF_BYTES + NXSLOT * xhat_bytes(NB8) + static_cast<int>(sizeof(KtvSmem));
  auto kern = ktv_kernel<NB8>;
  SAB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const long long nitems = static_cast<long long>((p.ni + p.chunk - 1) / p.chunk) * p.nj;
  const int grid = static_cast<int>(nitems < sm_count() ? nitems : sm_count());
  kern<<<grid, THREADS, smem, st>>>(p);
  SAB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int sa_ktv(void* stream, const void* kstore, int64_t a, int64_t b, const float* t, int dp, float* out,
                      void* workspace, int64_t workspace_bytes) {
  SAB_REQUIRE(kstore != nullptr && t != nullptr && out != nullptr && a > 0 && b > 0, "bad operands");
  SAB_REQUIRE((reinterpret_cast<uintptr_t>(kstore) & 127) == 0, "the kernel-block store must be 128-byte aligned");
  SAB_REQUIRE(dp >= 4 && dp <= 32 && dp % 4 == 0, "dp must be a multiple of 4 in [4, 32], got %d", dp);
  SAB_REQUIRE(workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & 127) == 0 &&
                  workspace_bytes >= sa_ktv_workspace_bytes(a),
              "workspace must be 128-byte aligned and hold sa_ktv_workspace_bytes(a) = %lld bytes",
              (long long)sa_ktv_workspace_bytes(a));
  KtvParams p;
  p.tiles = static_cast<const uint8_t*>(kstore);
  p.that = static_cast<const uint8_t*>(workspace);
  p.out = out;
  p.b = b;
  p.ni = static_cast<int>((a + TS - 1) / TS);
  p.nj = static_cast<int>((b + TS - 1) / TS);
  p.dp = dp;
  // row blocks per item: long enough to amortise the per-item atomics, short enough to balance the SMs
  int chunk = option_int("SAB_KTV_CHUNK", 0);
  if (chunk <= 0) {
    const long long want_items = 8ll * sm_count();
    chunk = static_cast<int>((static_cast<long long>(p.ni) * p.nj + want_items - 1) / want_items);
    if (chunk < 8) chunk = 8;
    if (chunk > 128) chunk = 128;
  }
  p.chunk = chunk < p.ni ? chunk : p.ni;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* that = static_cast<uint8_t*>(workspace);
  switch ((dp + 7) / 8) {
    case 1: return ktv_launch<1>(st, t, a, p, that);
    case 2: return ktv_launch<2>(st, t, a, p, that);
    case 3: return ktv_launch<3>(st, t, a, p, that);
    default: return ktv_launch<4>(st, t, a, p, that);
  }
}
