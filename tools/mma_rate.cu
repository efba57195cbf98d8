// Developer microbenchmark (not part of the library): SM cycles per tcgen05.mma kind::f16 as a function of
// the instruction shape, operand source (SS / TS) and CTA group.  Every CTA (pair) issues `reps` MMAs back
// to back from one thread on fixed shared-memory operands and times issue -> commit with clock64().
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o gpurun_out/mma_rate tools/mma_rate.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../sobolev_alignment_b200/csrc/ptx.cuh"

using namespace sab;

// nacc > 0: rotate the accumulator over `nacc` disjoint TMEM column ranges (independent chains).
// nacc < 0: "mix" mode -- per step 4 SS MMAs (N=224 -> columns [0,224)) followed by -nacc small TS MMAs
//           (shape N, accumulators at columns 448 / 480 alternating), the issue pattern of the fused kernel.
template <int CG, bool TS>
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int reps, int ksteps, int nacc, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, bar2;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  const int crank = CG > 1 ? static_cast<int>(cluster_ctarank()) : 0;
  // operands: A 128 x 64 fp16 (16 KB) at 0, B up to 256 x 64 fp16 (32 KB) at 16 KB, 128B swizzle
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u ^ (i * 2654435761u & 0x03ff03ffu);
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    mbar_init(smem_u32(&bar2), 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc<CG>(smem_u32(&slot), 512);
    tmem_relinquish<CG>();
  }
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  if (CG > 1) cluster_sync();
  tcgen05_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x == 32 && crank == 0) {
    const uint32_t idesc = umma_idesc_f16(128 * CG, N);
    const uint64_t da = umma_smem_desc_sw128(smem_u32(smem));
    const uint64_t db = umma_smem_desc_sw128(smem_u32(smem) + 16384);
    // warm-up
    for (int k = 0; k < 4; ++k) {
      if (TS) umma_f16_ts<CG>(tmem, tmem + 256, db + 2 * (k & 3), idesc, 1u);
      else umma_f16_ss<CG>(tmem, da + 2 * (k & 3), db + 2 * (k & 3), idesc, 1u);
    }
    umma_commit<CG>(smem_u32(&bar), 3);
    mbar_wait(smem_u32(&bar), 0);
    const long long t0 = clock64();
    if (nacc >= 1000) {
      // tile pattern of the fused kernel: per tile 23 x (8 MMAs + commit) + (4 MMAs + commit), then
      // `extra` more commits; bit 4 of the variant alternates the accumulator between two TMEM buffers
      const int variant = nacc - 1000;
      const int extra = variant & 3;
      const bool alternate = (variant & 4) != 0;
      const bool odd_tail = (variant & 8) == 0;
      const uint32_t idesc_big = umma_idesc_f16(128 * CG, 224);
      for (int r = 0; r < reps; ++r) {
        const uint32_t d = tmem + ((alternate && (r & 1)) ? 224u : 0u);
#pragma unroll 1
        for (int sidx = 0; sidx < 24; ++sidx) {
          const int n = (sidx == 23 && odd_tail) ? 4 : 8;
          for (int q = 0; q < n; ++q)
            umma_f16_ss<CG>(d, da + 2 * (q & 3), db + 2 * (q & 3), idesc_big, (sidx | q) ? 1u : 0u);
          umma_commit<CG>(smem_u32(&bar2), 3);
        }
        for (int e = 0; e < extra; ++e) umma_commit<CG>(smem_u32(&bar2), 3);
      }
    } else if (nacc >= 100) {
      // issue pattern of the fused kernel's main loop: groups of `per` MMAs, each group followed by a
      // tcgen05.commit to a scratch barrier (nacc in [100,200)) and preceded by an (always satisfied)
      // mbarrier wait + tcgen05 fence (nacc >= 200)
      const int per = nacc % 100;
      const uint32_t idesc_big = umma_idesc_f16(128 * CG, N);
      for (int r = 0; r < reps; ++r) {
#pragma unroll 1
        for (int k = 0; k < ksteps; k += per) {
          if (nacc >= 200) {
            mbar_wait(smem_u32(&bar), 0);
            tcgen05_fence_after();
          }
          for (int q = 0; q < per; ++q) umma_f16_ss<CG>(tmem, da + 2 * (q & 3), db + 2 * (q & 3), idesc_big, 1u);
          umma_commit<CG>(smem_u32(&bar2), 3);
        }
      }
    } else if (nacc < 0) {
      const uint32_t idesc_big = umma_idesc_f16(128 * CG, 224);
      for (int r = 0; r < reps; ++r) {
#pragma unroll 1
        for (int k = 0; k < ksteps; ++k) {
          for (int q = 0; q < 4; ++q) umma_f16_ss<CG>(tmem, da + 2 * q, db + 2 * q, idesc_big, 1u);
          for (int q = 0; q < -nacc; ++q)
            umma_f16_ts<CG>(tmem + 448 + 32 * (q & 1), tmem + 232 + 8 * q, db + 2 * (q & 3), idesc, 1u);
        }
      }
    } else {
      for (int r = 0; r < reps; ++r) {
#pragma unroll 1
        for (int k = 0; k < ksteps; ++k) {
          const uint32_t d = tmem + static_cast<uint32_t>((k % nacc) * N);
          if (TS) umma_f16_ts<CG>(d, tmem + 256, db + 2 * (k & 3), idesc, 1u);
          else umma_f16_ss<CG>(d, da + 2 * (k & 3), db + 2 * (k & 3), idesc, 1u);
        }
      }
    }
    umma_commit<CG>(smem_u32(&bar), 3);
    mbar_wait(smem_u32(&bar), 1);
    const long long t1 = clock64();
    cycles[blockIdx.x / CG] = t1 - t0;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (CG > 1) cluster_sync();
  if (warp == 0) {
    tcgen05_fence_after();
    tmem_dealloc<CG>(tmem, 512);
  }
}

template <int CG, bool TS>
static void run(int N, int grid_groups, int nacc = 1) {
  const int reps = 64, ksteps = 16;
  long long* d;
  cudaMalloc(&d, sizeof(long long) * grid_groups);
  auto kern = rate_kernel<CG, TS>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152 + 1024);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid_groups * CG);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = 49152 + 1024;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  for (int it = 0; it < 2; ++it) {
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, N, reps, ksteps, nacc, d);
    if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return; }
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); exit(1); }
  }
  std::vector<long long> h(grid_groups);
  cudaMemcpy(h.data(), d, sizeof(long long) * grid_groups, cudaMemcpyDeviceToHost);
  long long mn = h[0], mx = h[0];
  double sum = 0;
  for (auto v : h) { mn = v < mn ? v : mn; mx = v > mx ? v : mx; sum += v; }
  const double per = sum / grid_groups / (reps * ksteps);
  const double flops_per_sm_clk = 2.0 * 128 * N * 16 / per;
  if (nacc >= 1000) {
    const int v = nacc - 1000;
    const double mmas = (v & 8) ? 192.0 : 188.0;
    printf("CG=%d tile pattern: extra commits %d, alternate accumulators %d, odd tail %d: %.0f cycles/tile (ideal %.0f)\n",
           CG, v & 3, (v >> 2) & 1, (v & 8) ? 0 : 1, sum / grid_groups / 64.0, mmas * 112.4);
  } else if (nacc >= 100)
    printf("CG=%d SS N=%d, %s every %d MMAs: %.1f cycles/MMA\n", CG, N, nacc >= 200 ? "wait+fence+commit" : "commit", nacc % 100, per);
  else if (nacc < 0)
    printf("CG=%d mix: 4 x SS N=224 + %d x TS N=%d per step: %.1f cycles/step (4 x 112.4 = 449.6 alone)\n", CG, -nacc, N, per);
  else
    printf("CG=%d %s M=%d N=%3d nacc=%d groups=%3d: %.1f cycles/MMA (min %.1f max %.1f)  -> %.0f flops/clk/SM\n", CG,
           TS ? "TS" : "SS", 128 * CG, N, nacc, grid_groups, per, double(mn) / (reps * ksteps),
           double(mx) / (reps * ksteps), flops_per_sm_clk);
  cudaFree(d);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int groups = sms;
  for (int v : {0, 1, 2, 4, 6, 8, 12, 14}) run<2, false>(224, groups / 2, 1000 + v);
  return 0;
}
