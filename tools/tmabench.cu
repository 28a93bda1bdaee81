// Stand-alone experiment: how many bytes per clock can TMA deliver from L2 into shared memory, per SM and chip-wide, and what does
// cluster multicast change?  Every CTA runs the operand ring of the GEMM main loop without the MMAs: a producer thread issues the
// loads of one 32 KB stage (A-like 16 KB + B-like 16 KB, SWIZZLE_128B boxes of 64 bf16 columns) into a 6-deep ring, a consumer
// thread releases a stage as soon as it is full.  Source: a 32 MB bf16 matrix (L2-resident after the warm-up pass).
//   mode 0: unicast, 2 x 16 KB per stage                                  (tc_gemm_kernel today)
//   mode 1: cluster of 2, A unicast 16 KB + B as two 8 KB halves multicast to both CTAs   (L2 reads per CTA and stage: 24 KB)
//   mode 2: cluster of 2, A and B both as multicast halves                                  (16 KB)
//   mode 3: cluster of 4, A unicast + B as four 4 KB quarters multicast to all 4           (20 KB)
//   mode 4: cluster of 4, A and B both as multicast quarters                               (8 KB)
//   mode 5: unicast, only ONE 16 KB load per stage (half the traffic, same ring)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I instructany2pix_b200/csrc -o tools/tmabench tools/tmabench.cu
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>

#include "common.cuh"
using namespace ia2p;

constexpr int STAGES = 6, STAGE_BYTES = 32768;

__global__ void __launch_bounds__(64, 1) tma_ring(const __grid_constant__ CUtensorMap m128, const __grid_constant__ CUtensorMap m64,
                                                  const __grid_constant__ CUtensorMap m32, int mode, int iters, int csize,
                                                  long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = base + STAGES * STAGE_BYTES;
  auto full = [&](int s) { return bars + 8u * s; };
  auto empty = [&](int s) { return bars + 8u * (STAGES + s); };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = csize > 1 ? cluster_ctarank() : 0u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), csize); }
    fence_barrier_init();
  }
  if (csize > 1) cluster_sync_all(); else __syncthreads();
  const int cta = blockIdx.x, cl = cta / csize;
  const int rowA = (cta * 128) % 4096, rowB = (cl * 128 + 2048) % 4096;     // A: private rows; B: rows shared by the cluster
  if (warp == 0 && lane == 0) {
    int stage = 0; uint32_t phase = 0;
    const long long t0 = clock64();
    unsigned long long g0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    for (int i = 0; i < iters; ++i) {
      mbar_wait(empty(stage), phase ^ 1u);
      const uint32_t dst = base + stage * STAGE_BYTES;
      const int col = (i * 64) % 4096;
      const uint16_t mask = (uint16_t)((1u << csize) - 1u);
      mbar_arrive_expect_tx(full(stage), mode == 5 ? 16384 : 32768);
      if (mode == 0 || mode == 5) {
        tma_load_2d(dst, &m128, full(stage), col, rowA);
        if (mode == 0) tma_load_2d(dst + 16384, &m128, full(stage), col, rowB);
      } else {
        const CUtensorMap* mp = csize == 2 ? &m64 : &m32;
        const int part = 128 / csize, pbytes = 16384 / csize;
        const bool mcA = (mode == 2 || mode == 4);
        if (mcA) {
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
                       ::"r"(dst + crank * pbytes), "l"(mp), "r"(full(stage)), "r"(col), "r"(rowB + 1024 + (int)crank * part), "h"(mask) : "memory");
        } else {
          tma_load_2d(dst, &m128, full(stage), col, rowA);
        }
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
                     ::"r"(dst + 16384 + crank * pbytes), "l"(mp), "r"(full(stage)), "r"(col), "r"(rowB + (int)crank * part), "h"(mask) : "memory");
      }
      if (++stage == STAGES) { stage = 0; phase ^= 1u; }
    }
    // drain: wait until the last STAGES stages were consumed
    for (int k = 0; k < STAGES; ++k) { mbar_wait(empty(stage), phase ^ 1u); if (++stage == STAGES) { stage = 0; phase ^= 1u; } }
    unsigned long long g1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    out[2 * cta] = clock64() - t0;
    out[2 * cta + 1] = (long long)(g1 - g0);
  } else if (warp == 1 && lane == 0) {
    int stage = 0; uint32_t phase = 0;
    for (int i = 0; i < iters; ++i) {
      mbar_wait(full(stage), phase);
      if (csize == 1) mbar_arrive(empty(stage));
      else for (int r = 0; r < csize; ++r) mbar_arrive_cluster(mapa_shared(empty(stage), (uint32_t)r));
      if (++stage == STAGES) { stage = 0; phase ^= 1u; }
    }
  }
  if (csize > 1) cluster_sync_all(); else __syncthreads();
}

typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                          const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* f = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
  EncFn enc = (EncFn)f;
  const size_t R = 4096, C = 4096;
  void* buf; cudaMalloc(&buf, R * C * 2); cudaMemset(buf, 0, R * C * 2);
  CUtensorMap maps[3];
  const int rows[3] = {128, 64, 32};
  for (int i = 0; i < 3; ++i) {
    cuuint64_t gd[2] = {C, R}, gs[1] = {C * 2};
    cuuint32_t bx[2] = {64, (cuuint32_t)rows[i]}, es[2] = {1, 1};
    CUresult r = enc(&maps[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  }
  long long* d; cudaMalloc(&d, 2 * 148 * sizeof(long long));
  const int smem = STAGES * STAGE_BYTES + 1024 + 256;
  cudaFuncSetAttribute(tma_ring, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 4000;
  const int csz[6] = {1, 2, 2, 4, 4, 1};
  const double l2kb[6] = {32, 24, 16, 20, 8, 16};
  for (int grid : {148, 16})
    for (int mode = 0; mode < 6; ++mode) {
      int g = grid - grid % csz[mode];
      for (int rep = 0; rep < 2; ++rep) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(g); cfg.blockDim = dim3(64); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = csz[mode]; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, tma_ring, maps[0], maps[1], maps[2], mode, iters, csz[mode], d);
      }
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
      long long h[2 * 148];
      cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      double clk = 0, ns = 0;
      for (int i = 0; i < g; ++i) { clk += (double)h[2 * i]; if ((double)h[2 * i + 1] > ns) ns = (double)h[2 * i + 1]; }
      clk /= g;
      const double recv = (mode == 5 ? 16384.0 : 32768.0) * iters;
      printf("grid %3d mode %d (cluster %d): %7.1f clk per stage -> %5.1f B/clk/SM received; chip: received %6.2f TB/s, read from L2 %6.2f TB/s  (%.2f GHz)\n",
             g, mode, csz[mode], clk / iters, recv / clk, recv * g / ns / 1e3, l2kb[mode] * 1024.0 * iters * g / ns / 1e3, clk / ns);
    }
  return 0;
}
