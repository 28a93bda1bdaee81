// Stand-alone experiment: how long does one tcgen05.mma.kind::f16 take as a function of its N, of where A lives, of the CTA group
// and of the accumulator dependency chain?
// One CTA per SM, one thread issues `iters` groups of 4 MMAs (one 64-wide k-block: +32 B descriptor steps, like the GEMM main
// loop) on whatever bytes are in shared memory, then commits and waits; clock64 around the whole chain.
//   mode 0: A and B from shared memory (SWIZZLE_128B K-major), one accumulator       -- tc_gemm_kernel<*, 1, *>
//   mode 1: A from TMEM, B from shared memory (MN-major)                            -- the P.V MMA of fa_tc_kernel
//   mode 2/4: round-robin over 2/4 accumulators with the B rows split likewise (round 1)
//   mode 10: like 0, but consecutive MMAs alternate between TWO accumulators (same operands): is the per-instruction overhead a
//            read-after-write bubble on the accumulator?
//   mode 11: like 0, with a tcgen05.commit after every 4 MMAs (the main loop's stage release)
//   mode 20: cta_group::2, M = 256 over a CTA pair (each CTA: its 128 A rows + N/2 B rows), one accumulator -- tc_gemm_kernel<*, 2, *>
//   mode 21: cta_group::2, alternating between two accumulators
//   mode 22: cta_group::2, commit (multicast to both CTAs) after every 4 MMAs
//   mode 23: like 22, plus the main loop's per-k-block mbarrier wait (on a barrier whose phase completed long ago) + tcgen05 fence
//            BEFORE the 4 MMAs: does the issuing thread run far enough ahead of the tensor pipe to hide that latency?
//   mode 24: like 23, but the wait for the NEXT k-block sits between the 2nd and 3rd MMA of the current one (software pipelining)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I instructany2pix_b200/csrc -o tools/mmabench tools/mmabench.cu
#include <cstdio>
#include <cuda_runtime.h>

#include "common.cuh"
using namespace ia2p;

__global__ void __launch_bounds__(128, 1) mma_chain(int n, int mode, int iters, int ctas_active, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + 16384, bar = base + 16384 + 32768, slot = bar + 16, bar2 = bar + 8;
  const int warp = warp_idx_uniform(), lane = threadIdx.x & 31;
  (void)lane;
  if ((int)blockIdx.x >= ctas_active) return;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar2, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
  if (warp == 1 && elect_one()) {     // warp-uniform role + elect.sync lane: uniform-register operands (round 1 used lane == 0: R2UR waterfall, 158.8 clk flat)
    const uint32_t idesc = umma_idesc_bf16(128, n, mode == 1 ? 1 : 0);
    const uint64_t da = umma_desc_sw128(sA), db = umma_desc_sw128(sB);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (mode == 0 || mode == 11) umma_bf16(tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, 1u);
        else if (mode == 10) umma_bf16(tmem + (uint32_t)((k & 1) * 256), da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, 1u);
        else if (mode == 2 || mode == 4) {    // mode = number of independent accumulators, round-robin (B rows split likewise)
          for (int a = 0; a < mode; ++a)
            umma_bf16(tmem + (uint32_t)(a * n), da + (uint64_t)(2 * k), db + (uint64_t)(2 * k) + (uint64_t)((a * n * 128) >> 4), idesc, 1u);
        } else umma_bf16_ts(tmem, tmem + 256u + (uint32_t)(k * 8), umma_desc_sw128_mn(sB + k * 2048, 16384), idesc, 1u);
      }
      if (mode == 11) umma_commit(bar2);
    }
    umma_commit(bar);
    mbar_wait(bar, 0);
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// CTA pair: cluster of 2, the leader issues tcgen05.mma.cta_group::2 (M = 256, each CTA's TMEM receives its 128 rows)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) mma_chain_pair(int n, int mode, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + 16384, bar = base + 16384 + 32768, slot = bar + 32, bar2 = bar + 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar2, 1); mbar_init(bar + 24, 1); fence_barrier_init(); mbar_arrive(bar + 24); }
  if (warp == 0) { tmem_alloc_2sm(slot, 512); tmem_relinquish_2sm(); }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
  if (warp == 1 && lane == 0) {
    const long long t0 = clock64();
    if (rank == 0) {
      const uint32_t idesc = umma_idesc_bf16(256, n, 0);
      const uint64_t da = umma_desc_sw128(sA), db = umma_desc_sw128(sB);
      const uint32_t bar3 = bar + 24;                     // completes phase 0 once, then stays "already complete" for parity 0
      for (int i = 0; i < iters; ++i) {
        if (mode == 23) { mbar_wait(bar3, 0); tc_fence_after(); }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t d = (mode == 21) ? tmem + (uint32_t)((k & 1) * 256) : tmem;
          umma_bf16_2sm(d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, 1u);
          if (mode == 24 && k == 1) { mbar_wait(bar3, 0); tc_fence_after(); }
        }
        if (mode >= 22) umma_commit_2sm_mc(bar2, 3);
      }
      umma_commit_2sm_mc(bar, 3);
    }
    mbar_wait(bar, 0);
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 0) { tc_fence_after(); tmem_dealloc_2sm(tmem, 512); }
}


// The GEMM main loop minus TMA and epilogue: the leader walks a 6-stage ring of REAL (random bf16) operands, with the per-k-block
// wait + fence + 4 MMAs + commit sequence; optional spinning warps like the kernel's idle epilogue / producer warps.
//   mode 30: zeros in smem, one stage;  31: random data, one stage;  32: random data, 6-stage ring;  33: 32 + 8 warps spinning on
//   mbarrier.try_wait (the epilogue warps waiting for an accumulator);  34: 33 with the spin replaced by one long suspended wait
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(320, 1) mma_ring_pair(int mode, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar = base + 6 * 32768, bar2 = bar + 8, bar3 = bar + 24, slot = bar + 32, barN = bar + 40;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  // fill the ring
  uint32_t* sm = reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)));
  for (int i = threadIdx.x; i < 6 * 32768 / 4; i += blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 97u;
    h ^= h >> 13; h *= 0x5bd1e995u; h ^= h >> 15;
    // two bf16 in [-1, 1): exponent 0x3f (0.5..1) / random sign, random mantissa
    const uint32_t v = (mode == 30) ? 0u : ((h & 0x807f807fu) | 0x3f003f00u);
    sm[i] = v;
  }
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar2, 1); mbar_init(bar3, 1); mbar_init(barN, 1); fence_barrier_init(); mbar_arrive(bar3); }
  if (warp == 0) { tmem_alloc_2sm(slot, 512); tmem_relinquish_2sm(); }
  fence_proxy_async();
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
  if (warp == 1 && lane == 0) {
    const long long t0 = clock64();
    if (rank == 0) {
      const uint32_t idesc = umma_idesc_bf16(256, 256, 0);
      int stage = 0;
      for (int i = 0; i < iters; ++i) {
        mbar_wait(bar3, 0); tc_fence_after();
        const uint32_t a = base + (mode >= 32 ? stage * 32768 : 0);
        const uint64_t da = umma_desc_sw128(a), db = umma_desc_sw128(a + 16384);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_2sm(tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (uint32_t)((i | k) != 0));
        umma_commit_2sm_mc(bar2, 3);
        if (++stage == 6) stage = 0;
      }
      umma_commit_2sm_mc(bar, 3);
    }
    mbar_wait(bar, 0);
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
    mbar_arrive(barN);                                     // release the spinning warps
  } else if (warp >= 2 && mode >= 33) {
    if (mode == 33) { while (!mbar_try_wait(barN, 0)) { if (clock64() < 0) __trap(); } }
    else {
      uint32_t ok = 0;
      while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                               : "=r"(ok) : "r"(barN), "r"(0u), "r"(0x989680u) : "memory");
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 0) { tc_fence_after(); tmem_dealloc_2sm(tmem, 512); }
}

// The flash-attention kernel's tensor work per 128-key block and CTA, issued back to back with compile-time shapes (no run-time
// mode branches in the issue loop): WHAT = 1: S = Q K^T as 4 x (M128 N128 K16, A smem); 2: O += P V as 8 x (M128 N64 K16, A tmem,
// MN-major B); 3: both; 4: P V as N 128 (two heads' worth of V columns -- what a wider N would cost).  One or two CTAs per SM.
template <int WHAT>
__global__ void __launch_bounds__(128, 2) mma_fa_pattern(int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + 16384, bar = base + 16384 + 32768, slot = bar + 16;
  const int warp = warp_idx_uniform();
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(slot, 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
  if (warp == 1 && elect_one()) {
    constexpr uint32_t idesc_qk = umma_idesc_bf16(128, 128, 0);
    constexpr uint32_t idesc_pv = umma_idesc_bf16(128, WHAT == 4 ? 128 : 64, 1);
    const uint64_t da = umma_desc_sw128(sA), db = umma_desc_sw128(sB);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if (WHAT & 1) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc_qk, 1u);
      }
      if (WHAT & 6) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          umma_bf16_ts(tmem + 128u, tmem + 192u + (uint32_t)(ks * 8), umma_desc_sw128_mn(sB + ks * 2048, 16384), idesc_pv, 1u);
      }
    }
    umma_commit(bar);
    mbar_wait(bar, 0);
    out[blockIdx.x] = clock64() - t0;
  }
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

template <int WHAT>
static void run_fa_pattern(long long* d, int smem, const char* what) {
  cudaFuncSetAttribute(mma_fa_pattern<WHAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int per_sm : {1, 2}) {
    const int ctas = 148 * per_sm, iters = 2000;
    for (int rep = 0; rep < 2; ++rep) mma_fa_pattern<WHAT><<<ctas, 128, smem>>>(iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("fa pattern %d: %s\n", WHAT, cudaGetErrorString(e)); return; }
    long long h[296];
    cudaMemcpy(h, d, ctas * sizeof(long long), cudaMemcpyDeviceToHost);
    double s = 0;
    for (int i = 0; i < ctas; ++i) s += (double)h[i];
    printf("fa pattern %-44s %d CTA/SM: %7.1f clk per 128-key block and CTA\n", what, per_sm, s / ctas / iters);
  }
}

int main() {
  long long* d;
  cudaMalloc(&d, 296 * sizeof(long long));
  const int smem = 16384 + 32768 + 1024 + 64;
  run_fa_pattern<1>(d, smem, "QK^T: 4 x M128 N128 K16");
  run_fa_pattern<2>(d, smem, "PV: 8 x M128 N64 K16 (A in TMEM)");
  run_fa_pattern<3>(d, smem, "QK^T + PV");
  run_fa_pattern<4>(d, smem, "PV at N128: 8 x M128 N128 K16 (A in TMEM)");
  cudaFuncSetAttribute(mma_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(mma_chain_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 2000;
  for (int mode : {0, 1, 2, 4, 10, 11})
    for (int ctas : {148})
      for (int n : {16, 32, 64, 96, 128, 160, 192, 256}) {
        if (mode == 1 && n > 128) continue;                 // MN-major B of 64-element atoms: keep to what the kernel uses
        if ((mode == 2 || mode == 4) && (mode * n > 512 || n % 64 != 0 || mode * n * 128 > 32768)) continue;
        if (mode >= 10 && n < 64) continue;
        for (int rep = 0; rep < 2; ++rep) mma_chain<<<148, 128, smem>>>(n, mode, iters, ctas, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d n %d: %s\n", mode, n, cudaGetErrorString(e)); return 1; }
        long long h[148];
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        double s = 0;
        for (int i = 0; i < ctas; ++i) s += (double)h[i];
        const double clk = s / ctas / (iters * 4.0 * ((mode == 2 || mode == 4) ? mode : 1));
        printf("mode %2d (%s) CTAs %3d  M128 N%-3d K16: %6.1f clk per MMA  -> %6.0f MAC/clk/SM (peak 4096)\n", mode,
               mode == 0 ? "A smem" : mode == 1 ? "A tmem" : mode == 10 ? "A smem, 2 alternating accumulators" :
               mode == 11 ? "A smem, commit per 4" : "A smem, round-robin accumulators", ctas, n, clk, 128.0 * n * 16 / clk);
      }
  for (int mode : {20, 22, 23, 24})
    for (int n : {64, 128, 160, 192, 256}) {
      for (int rep = 0; rep < 2; ++rep) mma_chain_pair<<<148, 128, smem>>>(n, mode, iters, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d n %d: %s\n", mode, n, cudaGetErrorString(e)); return 1; }
      long long h[148];
      cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      double s = 0;
      for (int i = 0; i < 148; i += 2) s += (double)h[i];   // leader CTAs
      const double clk = s / 74 / (iters * 4.0);
      printf("mode %2d (cta_group::2%s) pairs 74  M256 N%-3d K16: %6.1f clk per MMA  -> %6.0f MAC/clk/SM (peak 4096)\n", mode,
             mode == 21 ? ", 2 alternating accumulators" : mode == 22 ? ", commit per 4" : mode == 23 ? ", wait+fence then 4 MMAs, commit" : mode == 24 ? ", wait+fence between MMA 2 and 3, commit" : "", n, clk, 128.0 * n * 16 / clk);
    }
  {
    const int smem2 = 6 * 32768 + 1024 + 128;
    cudaFuncSetAttribute(mma_ring_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2);
    for (int mode : {30, 31, 32, 33, 34}) {
      for (int rep = 0; rep < 2; ++rep) mma_ring_pair<<<148, 320, smem2>>>(mode, 4000, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("ring mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
      long long h[148];
      cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      double s = 0;
      for (int i = 0; i < 148; i += 2) s += (double)h[i];
      printf("ring mode %d (%s): %6.1f clk per k-block of 4 MMAs (M256 N256; ideal 512)\n", mode,
             mode == 30 ? "zeros, one stage" : mode == 31 ? "random data, one stage" : mode == 32 ? "random data, 6-stage ring" :
             mode == 33 ? "ring + 8 warps spinning on try_wait" : "ring + 8 warps in one long try_wait", s / 74 / 4000);
    }
  }
  return 0;
}
