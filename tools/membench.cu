// Stand-alone experiment: DRAM efficiency of the GEMM epilogue's access patterns for out[M,N] = res[M,N] + c (fp32).
//   A: thread = row, 32-column chunks, 256-bit accesses (what tc_gemm_kernel does: 32 rows x 1 sector per instruction)
//   B: fully coalesced (consecutive lanes -> consecutive 16 B)
//   C: thread = row but with a 4-lane transpose so each instruction covers 8 rows x 4 sectors (whole 128-byte lines)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o membench tools/membench.cu ; run: ./membench
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void ldg256(const float* p, float* d) {
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3]), "=f"(d[4]), "=f"(d[5]), "=f"(d[6]), "=f"(d[7]) : "l"(p));
}
__device__ __forceinline__ void stg256(float* p, const float* d) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"l"(p), "f"(d[0]), "f"(d[1]), "f"(d[2]), "f"(d[3]), "f"(d[4]), "f"(d[5]), "f"(d[6]), "f"(d[7]) : "memory");
}

// A: grid = (M/128) * (N/256) tiles like the GEMM; 256 threads = 8 warps: quarter q = w&3, half = w>>2, 4 chunks each
// A2: like A, mode 1 = loads only (result folded into one predicated store), mode 2 = stores only
__global__ void patA2(const float* __restrict__ res, float* __restrict__ out, int M, int N, int tiles_n, int mode) {
  for (int tile = blockIdx.x; tile < (M / 128) * tiles_n; tile += gridDim.x) {
    const int mt = tile / tiles_n, nt = tile % tiles_n;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long row = mt * 128 + (w & 3) * 32 + lane;
    const int c0 = nt * 256 + (w >> 2) * 128;
    float r[4][32];
    if (mode != 2) {
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int i = 0; i < 4; ++i) ldg256(res + row * N + c0 + c * 32 + i * 8, &r[c][i * 8]);
    } else {
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int i = 0; i < 32; ++i) r[c][i] = (float)(tile + i);
    }
    if (mode != 1) {
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int i = 0; i < 4; ++i) stg256(out + row * N + c0 + c * 32 + i * 8, &r[c][i * 8]);
    } else {
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int i = 0; i < 32; ++i) s += r[c][i];
      if (s == 12345.678f) out[row] = s;
    }
  }
}
__global__ void patA(const float* __restrict__ res, float* __restrict__ out, int M, int N, int tiles_n) {
  for (int tile = blockIdx.x; tile < (M / 128) * tiles_n; tile += gridDim.x) {
    const int mt = tile / tiles_n, nt = tile % tiles_n;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long row = mt * 128 + (w & 3) * 32 + lane;
    const int c0 = nt * 256 + (w >> 2) * 128;
    float r[4][32];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int i = 0; i < 4; ++i) ldg256(res + row * N + c0 + c * 32 + i * 8, &r[c][i * 8]);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
#pragma unroll
      for (int i = 0; i < 32; ++i) r[c][i] += 1.0f;
#pragma unroll
      for (int i = 0; i < 4; ++i) stg256(out + row * N + c0 + c * 32 + i * 8, &r[c][i * 8]);
    }
  }
}
// B: coalesced float4 grid-stride
__global__ void patB(const float4* __restrict__ res, float4* __restrict__ out, long long n4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 v = res[i];
    v.x += 1.f; v.y += 1.f; v.z += 1.f; v.w += 1.f;
    out[i] = v;
  }
}
// C: like A but lanes regrouped so one instruction touches 8 rows x 4 sectors (row = 8*j + lane/4, sector = lane%4)
__global__ void patC(const float* __restrict__ res, float* __restrict__ out, int M, int N, int tiles_n) {
  for (int tile = blockIdx.x; tile < (M / 128) * tiles_n; tile += gridDim.x) {
    const int mt = tile / tiles_n, nt = tile % tiles_n;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long row0 = mt * 128 + (w & 3) * 32;
    const int c0 = nt * 256 + (w >> 2) * 128;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float r[4][8];
#pragma unroll
      for (int j = 0; j < 4; ++j) ldg256(res + (row0 + j * 8 + (lane >> 2)) * N + c0 + c * 32 + (lane & 3) * 8, r[j]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int i = 0; i < 8; ++i) r[j][i] += 1.0f;
        stg256(out + (row0 + j * 8 + (lane >> 2)) * N + c0 + c * 32 + (lane & 3) * 8, r[j]);
      }
    }
  }
}

int main() {
  const int M = 8192, N = 1280, NB = 8;
  float *res, *out;
  cudaMalloc(&res, (size_t)NB * M * N * 4);
  cudaMalloc(&out, (size_t)NB * M * N * 4);
  cudaMemset(res, 0, (size_t)NB * M * N * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int pat = 0; pat < 3; ++pat) {
    for (int it = 0; it < 3; ++it) {
      if (pat == 0) patA<<<148, 256>>>(res, out, M, N, N / 256);
      if (pat == 1) patB<<<148 * 8, 256>>>((const float4*)res, (float4*)out, (long long)M * N / 4);
      if (pat == 2) patC<<<148, 256>>>(res, out, M, N, N / 256);
    }
    cudaEventRecord(e0);
    const int iters = 40;
    for (int it = 0; it < iters; ++it) {
      const float* r = res + (size_t)(it % NB) * M * N;
      float* o = out + (size_t)(it % NB) * M * N;
      if (pat == 0) patA<<<148, 256>>>(r, o, M, N, N / 256);
      if (pat == 1) patB<<<148 * 8, 256>>>((const float4*)r, (float4*)o, (long long)M * N / 4);
      if (pat == 2) patC<<<148, 256>>>(r, o, M, N, N / 256);
    }
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("pattern %c: %.1f us per pass, %.0f GB/s (read+write %d MB)  err=%s\n", "ABC"[pat], ms / iters * 1e3,
           2.0 * M * N * 4 / (ms / iters * 1e-3) / 1e9, (int)(2.0 * M * N * 4 / 1e6), cudaGetErrorString(cudaGetLastError()));
  }
  for (int nb = 8; nb >= 1; nb /= 8)
    for (int mode = 0; mode < 3; ++mode) {
      for (int it = 0; it < 3; ++it) patA2<<<148, 256>>>(res, out, M, N, N / 256, mode);
      cudaEventRecord(e0);
      const int iters = 40;
      for (int it = 0; it < iters; ++it)
        patA2<<<148, 256>>>(res + (size_t)(it % nb) * M * N, out + (size_t)(it % nb) * M * N, M, N, N / 256, mode);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      printf("pattern A2 buffers=%d mode=%s: %.1f us per pass, %.0f GB/s\n", nb, mode == 0 ? "rw" : mode == 1 ? "read" : "write",
             ms / iters * 1e3, (mode == 0 ? 2.0 : 1.0) * M * N * 4 / (ms / iters * 1e-3) / 1e9);
    }
  return 0;
}
