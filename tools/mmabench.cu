// Stand-alone experiment: how long does one tcgen05.mma.kind::f16 take as a function of its N (and of where A lives)?
// One CTA per SM, one thread issues `iters` groups of 4 MMAs (one 64-wide k-block: +32 B descriptor steps, like the GEMM main
// loop) on whatever bytes are in shared memory, then commits and waits; clock64 around the whole chain.
//   mode 0: A and B from shared memory (SWIZZLE_128B K-major)        -- tc_gemm_kernel
//   mode 1: A from TMEM, B from shared memory (MN-major)             -- the P.V MMA of fa_tc_kernel
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I instructany2pix_b200/csrc -o tools/mmabench tools/mmabench.cu
#include <cstdio>
#include <cuda_runtime.h>

#include "common.cuh"
using namespace ia2p;

__global__ void __launch_bounds__(128, 1) mma_chain(int n, int mode, int iters, int ctas_active, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + 16384, bar = base + 16384 + 32768, slot = bar + 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if ((int)blockIdx.x >= ctas_active) return;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
  if (warp == 1 && lane == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, n, mode == 1 ? 1 : 0);
    const uint64_t da = umma_desc_sw128(sA), db = umma_desc_sw128(sB);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (mode == 0) umma_bf16(tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, 1u);
        else if (mode >= 2) {                 // mode = number of independent accumulators, round-robin (B rows split likewise)
          for (int a = 0; a < mode; ++a)
            umma_bf16(tmem + (uint32_t)(a * n), da + (uint64_t)(2 * k), db + (uint64_t)(2 * k) + (uint64_t)((a * n * 128) >> 4), idesc, 1u);
        } else umma_bf16_ts(tmem, tmem + 256u + (uint32_t)(k * 8), umma_desc_sw128_mn(sB + k * 2048, 16384), idesc, 1u);
      }
    }
    umma_commit(bar);
    mbar_wait(bar, 0);
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  const int smem = 16384 + 32768 + 1024 + 64;
  cudaFuncSetAttribute(mma_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 2000;
  for (int mode : {0, 1, 2, 4})
    for (int ctas : {148})
      for (int n : {16, 32, 64, 96, 128, 160, 192, 256}) {
        if (mode == 1 && n > 128) continue;                 // MN-major B of 64-element atoms: keep to what the kernel uses
        if (mode >= 2 && (mode * n > 512 || n % 64 != 0 || mode * n * 128 > 32768)) continue;
        for (int rep = 0; rep < 2; ++rep) mma_chain<<<148, 128, smem>>>(n, mode, iters, ctas, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d n %d: %s\n", mode, n, cudaGetErrorString(e)); return 1; }
        long long h[148];
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        double s = 0;
        for (int i = 0; i < ctas; ++i) s += (double)h[i];
        const double clk = s / ctas / (iters * 4.0 * (mode >= 2 ? mode : 1));
        printf("mode %d (%s) CTAs %3d  M128 N%-3d K16: %6.1f clk per MMA  -> %6.0f MAC/clk/SM (peak 4096)\n", mode,
               mode == 0 ? "A smem" : mode == 1 ? "A tmem" : "A smem, round-robin accumulators", ctas, n, clk, 128.0 * n * 16 / clk);
      }
  return 0;
}
