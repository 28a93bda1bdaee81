// Issue / pipe micro-benchmark behind the flash-attention softmax (sm_100a): how many cycles per SM sub-partition a mix of
// MUFU.EX2, FFMA2 / FADD2 (packed fp32 pairs), scalar FFMA, F2FP (bf16x2 pack) and integer shift-adds costs with 1, 2 or 4
// resident warps per sub-partition.  Every stream is 8 independent chains, so latency is hidden and the numbers are throughput.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --expt-relaxed-constexpr -o tools/sfubench tools/sfubench.cu && tools/sfubench
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float ffma(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ uint32_t f2fp(float a, float b) { uint32_t d; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ uint32_t shadd(uint32_t a, uint32_t b) { uint32_t d; asm volatile("{.reg .u32 t; shl.b32 t, %1, 23; add.u32 %0, t, %2;}" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ float fmx(float a, float b) { float d; asm volatile("max.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }

constexpr int cmax(int a, int b) { return a > b ? a : b; }
// per loop iteration: NM MUFU, NF2 FFMA2, NF scalar FFMA, NP F2FP, NI shift-adds, NX FMNMX -- each on 8 rotating independent chains
template <int NM, int NF2, int NF, int NP, int NI, int NX>
__global__ void mix_kernel(long long* out, float* sink, int iters) {
  float m[8], f[8], x[8];
  uint64_t d[8];
  uint32_t q[8], n[8];
  for (int i = 0; i < 8; ++i) {
    m[i] = -0.001f * (threadIdx.x + i); f[i] = 0.5f + i; x[i] = 1.f + i;
    d[i] = ((uint64_t)__float_as_uint(1.f + i) << 32) | __float_as_uint(0.25f * i);
    q[i] = i; n[i] = i + threadIdx.x;
  }
  const uint64_t c2 = ((uint64_t)__float_as_uint(0.999f) << 32) | __float_as_uint(1.001f);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    // slot loop: the types are interleaved in proportion (Bresenham), consecutive instructions of a type use different chains
    constexpr int T = cmax(cmax(cmax(NM, NF2), cmax(NF, NP)), cmax(NI, NX));
#pragma unroll
    for (int s = 0; s < T; ++s) {
      if ((s + 1) * NM / T != s * NM / T) { const int r = (s * NM / T) & 7; m[r] = ex2(m[r]); }
      if ((s + 1) * NF2 / T != s * NF2 / T) { const int r = (s * NF2 / T) & 7; d[r] = fma2(d[r], c2, c2); }
      if ((s + 1) * NF / T != s * NF / T) { const int r = (s * NF / T) & 7; f[r] = ffma(f[r], 0.999f, 0.001f); }
      if ((s + 1) * NP / T != s * NP / T) { const int r = (s * NP / T) & 7; q[r] = f2fp(__uint_as_float(q[r]), f[r]); }
      if ((s + 1) * NI / T != s * NI / T) { const int r = (s * NI / T) & 7; n[r] = shadd(n[r], q[r]); }
      if ((s + 1) * NX / T != s * NX / T) { const int r = (s * NX / T) & 7; x[r] = fmx(x[r], -125.f); }
    }
  }
  const long long t1 = clock64();
  float acc = 0.f;
  for (int i = 0; i < 8; ++i) acc += m[i] + f[i] + x[i] + __uint_as_float((uint32_t)d[i]) + __uint_as_float(q[i]) + __uint_as_float(n[i]);
  if (acc == 123.456f) sink[0] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

template <int NM, int NF2, int NF, int NP, int NI, int NX>
static void run(const char* what) {
  long long* out; float* sink;
  cudaMalloc(&out, 8); cudaMalloc(&sink, 4);
  const int iters = 2000;
  printf("%-58s", what);
  for (int wps : {1, 2, 4}) {                          // warps per sub-partition (block = 4 sub-partitions x wps warps), one block per SM
    mix_kernel<NM, NF2, NF, NP, NI, NX><<<148, 128 * wps>>>(out, sink, iters);
    long long c = 0;
    cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
    printf("  %d w/SMSP: %7.1f clk/iter/warp-set", wps, (double)c / iters);
  }
  printf("\n");
  cudaFree(out); cudaFree(sink);
}

// legacy warp-level tensor-core path: mma.sync.m16n8k16 bf16 (what gemm_smallm_kernel / prior_trunk_kernel use): NCH independent accumulator chains
template <int NCH>
__global__ void hmma_kernel(long long* out, float* sink, int iters) {
  float c[NCH][4];
  for (int i = 0; i < NCH; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
  const unsigned a0 = 0x3f803f80u + threadIdx.x, a1 = 0x3f003f00u, a2 = 0x3e803e80u, a3 = 0x3f803f00u, b0 = 0x3f803f80u, b1 = 0x3f003f80u;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NCH; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  const long long t1 = clock64();
  float acc = 0.f;
  for (int i = 0; i < NCH; ++i) acc += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (acc == 123.456f) sink[0] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}
template <int NCH>
static void run_hmma() {
  long long* out; float* sink;
  cudaMalloc(&out, 8); cudaMalloc(&sink, 4);
  const int iters = 2000;
  printf("mma.sync m16n8k16 bf16, %d independent accumulators per warp:", NCH);
  for (int wps : {1, 2, 4}) {
    hmma_kernel<NCH><<<148, 128 * wps>>>(out, sink, iters);
    long long c = 0;
    cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
    printf("  %d w/SMSP: %6.1f clk per HMMA per sub-partition", wps, (double)c / iters / NCH / wps);
  }
  printf("\n");
  cudaFree(out); cudaFree(sink);
}

int main() {
  run_hmma<1>();
  run_hmma<2>();
  run_hmma<8>();
  printf("cycles per loop iteration (all resident warps run the same iteration concurrently)\n");
  run<8, 0, 0, 0, 0, 0>("8 MUFU.EX2");
  run<0, 8, 0, 0, 0, 0>("8 FFMA2");
  run<0, 0, 8, 0, 0, 0>("8 FFMA");
  run<0, 0, 0, 8, 0, 0>("8 F2FP");
  run<0, 0, 0, 0, 8, 0>("8 SHL+IADD (LEA)");
  run<0, 0, 0, 0, 0, 8>("8 FMNMX");
  run<8, 8, 0, 0, 0, 0>("8 MUFU + 8 FFMA2");
  run<8, 16, 0, 0, 0, 0>("8 MUFU + 16 FFMA2");
  run<8, 32, 0, 0, 0, 0>("8 MUFU + 32 FFMA2");
  run<8, 0, 32, 0, 0, 0>("8 MUFU + 32 FFMA");
  run<8, 0, 56, 0, 0, 0>("8 MUFU + 56 FFMA");
  run<8, 0, 0, 8, 0, 0>("8 MUFU + 8 F2FP");
  run<0, 16, 0, 8, 8, 8>("16 FFMA2 + 8 F2FP + 8 LEA + 8 FMNMX");
  // the softmax mixes per 16 key pairs (32 exponentials): POLY16 = 0, 4, 6, 8 (FFMA2 column counts FADD2 too)
  run<32, 32, 0, 16, 0, 0>("softmax mix POLY16=0: 32 MUFU 32 F2 16 F2FP");
  run<24, 56, 0, 16, 8, 8>("softmax mix POLY16=4: 24 MUFU 56 F2 16 F2FP 8 LEA 8 FMNMX");
  run<20, 68, 0, 16, 12, 12>("softmax mix POLY16=6: 20 MUFU 68 F2 16 F2FP 12 LEA 12 FMNMX");
  run<16, 80, 0, 16, 16, 16>("softmax mix POLY16=8: 16 MUFU 80 F2 16 F2FP 16 LEA 16 FMNMX");
  return 0;
}
