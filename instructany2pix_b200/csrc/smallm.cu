// Prior / embedding path: small-M weight-streaming GEMM and short-sequence causal attention.
//
// ia2p_gemm_smallm: out[M,N] = act(act_in(A) @ W^T + bias) + residual with fp32 activations and bf16 weights.
// M is tiny (<= 32 per pass: 2 x 11|14 GPT-2 tokens, or 2B embedding rows) so the op is bound by streaming W from
// HBM once.  Activations stay fp32-accurate on tensor cores by splitting A = hi + lo (two bf16 terms, ~16 mantissa
// bits) and issuing two warp-level mma.sync per k-step.  Each quad thread reads 16 contiguous bytes of a weight
// row (the k index is permuted identically for A and B, which leaves the dot product unchanged), so every 32-byte
// sector fetched is fully used.
#include "common.cuh"

namespace ia2p {

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void split_pair(float x, float y, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat16 xh = __float2bfloat16_rn(x), yh = __float2bfloat16_rn(y);
  const float xr = x - __bfloat162float(xh), yr = y - __bfloat162float(yh);
  __nv_bfloat162 h2; h2.x = xh; h2.y = yh;
  hi = *reinterpret_cast<uint32_t*>(&h2);
  lo = pack_bf16x2(xr, yr);
}

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == IA2P_ACT_GELU_NEW) return gelu_new_f(v);
  if (act == IA2P_ACT_SILU) return silu_f(v);
  return v;
}

// grid (ceil(N/32), ceil(M/32)); 128 threads: warp w -> output columns [32*bx + 8w, +8), rows [32*by, +32)
__global__ void __launch_bounds__(128)
gemm_smallm_kernel(const float* __restrict__ A, long long lda, const __nv_bfloat16* __restrict__ W,
                   const float* __restrict__ bias, const float* __restrict__ residual, long long ldr,
                   float* __restrict__ out, long long ldo, int M, int N, int K, int act_in, int act) {
  pdl_launch_dependents();   // programmatic dependent launch: see common.cuh
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int n0 = blockIdx.x * 32 + warp * 8;
  if (n0 >= N) return;
  const int m0 = blockIdx.y * 32;
  const __nv_bfloat16* wrow = W + (long long)(n0 + g) * K + t * 8;
  const float* arow[4];
  bool aok[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int row = m0 + g + 8 * r;            // r = 0,1 -> m-tile 0 rows g, g+8 ; r = 2,3 -> m-tile 1
    aok[r] = row < M;
    arow[r] = A + (long long)(aok[r] ? row : 0) * lda + t * 8;
  }
  float acc[2][4];
#pragma unroll
  for (int i = 0; i < 2; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;

#pragma unroll 4
  for (int k0 = 0; k0 < K; k0 += 32) {
    const uint4 wv = __ldg(reinterpret_cast<const uint4*>(wrow + k0));
    uint32_t ah[4][4], al[4][4];   // [row r][pair p]: pairs (x0,x1) (x2,x3) (x4,x5) (x6,x7)
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      float4 x0 = make_float4(0.f, 0.f, 0.f, 0.f), x1 = x0;
      if (aok[r]) {
        x0 = __ldg(reinterpret_cast<const float4*>(arow[r] + k0));
        x1 = __ldg(reinterpret_cast<const float4*>(arow[r] + k0) + 1);
        if (act_in == IA2P_ACT_SILU) {
          x0.x = silu_f(x0.x); x0.y = silu_f(x0.y); x0.z = silu_f(x0.z); x0.w = silu_f(x0.w);
          x1.x = silu_f(x1.x); x1.y = silu_f(x1.y); x1.z = silu_f(x1.z); x1.w = silu_f(x1.w);
        }
      }
      split_pair(x0.x, x0.y, ah[r][0], al[r][0]);
      split_pair(x0.z, x0.w, ah[r][1], al[r][1]);
      split_pair(x1.x, x1.y, ah[r][2], al[r][2]);
      split_pair(x1.z, x1.w, ah[r][3], al[r][3]);
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      // k-step A: slots (2t,2t+1)->pair0, (2t+8,2t+9)->pair1 ; k-step B: pair2, pair3.  a0/a2 = row g, a1/a3 = row g+8
      mma_bf16_16816(acc[mt], ah[2 * mt][0], ah[2 * mt + 1][0], ah[2 * mt][1], ah[2 * mt + 1][1], wv.x, wv.y);
      mma_bf16_16816(acc[mt], al[2 * mt][0], al[2 * mt + 1][0], al[2 * mt][1], al[2 * mt + 1][1], wv.x, wv.y);
      mma_bf16_16816(acc[mt], ah[2 * mt][2], ah[2 * mt + 1][2], ah[2 * mt][3], ah[2 * mt + 1][3], wv.z, wv.w);
      mma_bf16_16816(acc[mt], al[2 * mt][2], al[2 * mt + 1][2], al[2 * mt][3], al[2 * mt + 1][3], wv.z, wv.w);
    }
  }
  const int n = n0 + 2 * t;
  const float b0 = bias ? bias[n] : 0.f, b1 = bias ? bias[n + 1] : 0.f;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      const int row = m0 + 16 * mt + 8 * hf + g;
      if (row >= M) continue;
      float v0 = apply_act(acc[mt][2 * hf] + b0, act), v1 = apply_act(acc[mt][2 * hf + 1] + b1, act);
      if (residual) { v0 += residual[(long long)row * ldr + n]; v1 += residual[(long long)row * ldr + n + 1]; }
      *reinterpret_cast<float2*>(out + (long long)row * ldo + n) = make_float2(v0, v1);
    }
}

// grid (heads, batch); block (32, T): warp i = query i, lane l = head dims (2l, 2l+1)
__global__ void causal_attn_small_kernel(const float* __restrict__ qkv, float* __restrict__ out, int T, int E) {
  pdl_launch_dependents();   // programmatic dependent launch: see common.cuh
  pdl_wait();
  const int i = threadIdx.y, lane = threadIdx.x;
  const int h = blockIdx.x;
  const long long b = blockIdx.y;
  const float* base = qkv + b * T * 3LL * E + h * 64 + 2 * lane;
  const float2 qv = *reinterpret_cast<const float2*>(base + (long long)i * 3 * E);
  float sc[32];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    sc[j] = -INFINITY;
    if (j <= i) {                              // warp-uniform
      const float2 kv = *reinterpret_cast<const float2*>(base + (long long)j * 3 * E + E);
      sc[j] = warp_sum(qv.x * kv.x + qv.y * kv.y) * 0.125f;
      mx = fmaxf(mx, sc[j]);
    }
  }
  float sum = 0.f, o0 = 0.f, o1 = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    if (j <= i) {
      const float p = __expf(sc[j] - mx);
      sum += p;
      const float2 vv = *reinterpret_cast<const float2*>(base + (long long)j * 3 * E + 2 * E);
      o0 += p * vv.x;
      o1 += p * vv.y;
    }
  }
  *reinterpret_cast<float2*>(out + (b * T + i) * (long long)E + h * 64 + 2 * lane) = make_float2(o0 / sum, o1 / sum);
}

}  // namespace ia2p

using namespace ia2p;

extern "C" int ia2p_gemm_smallm(const float* A, int64_t lda, const void* W, const float* bias, const float* residual,
                                int64_t ldr, float* out, int64_t ldo, int64_t M, int64_t N, int64_t K, int act_in, int act,
                                void* stream) {
  if (int e = check_device()) return e;
  IA2P_REQUIRE(A && W && out && M > 0 && N > 0 && K > 0, IA2P_E_ARG, "gemm_smallm: bad arguments");
  IA2P_REQUIRE(K % 32 == 0 && N % 8 == 0, IA2P_E_SHAPE, "gemm_smallm: K=%lld must be a multiple of 32, N=%lld of 8", (long long)K, (long long)N);
  IA2P_REQUIRE(lda % 4 == 0 && ldo % 2 == 0, IA2P_E_ALIGN, "gemm_smallm: lda%%4, ldo%%2 required");
  IA2P_REQUIRE(act_in == IA2P_ACT_NONE || act_in == IA2P_ACT_SILU, IA2P_E_ARG, "gemm_smallm: act_in must be none or silu");
  const dim3 grid((unsigned)((N + 31) / 32), (unsigned)((M + 31) / 32));
  gemm_smallm_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(A, lda, static_cast<const __nv_bfloat16*>(W), bias,
                                                                          residual, ldr, out, ldo, (int)M, (int)N, (int)K,
                                                                          act_in, act);
  IA2P_LAUNCH_CHECK();
  return 0;
}

extern "C" int ia2p_causal_attn_small_f32(const float* qkv, float* out, int64_t batch, int64_t T, int heads, void* stream) {
  if (int e = check_device()) return e;
  IA2P_REQUIRE(qkv && out && batch > 0 && heads > 0, IA2P_E_ARG, "causal_attn_small: bad arguments");
  IA2P_REQUIRE(T >= 1 && T <= 32, IA2P_E_SHAPE, "causal_attn_small: T=%lld must be in [1,32]", (long long)T);
  const dim3 grid((unsigned)heads, (unsigned)batch), block(32, (unsigned)T);
  causal_attn_small_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(qkv, out, (int)T, heads * 64);
  IA2P_LAUNCH_CHECK();
  return 0;
}
