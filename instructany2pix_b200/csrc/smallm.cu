// Prior / embedding path: small-M weight-streaming GEMM and short-sequence causal attention.
//
// ia2p_gemm_smallm: out[M,N] = act(act_in(A) @ W^T + bias) + residual with fp32 activations and bf16 weights.
// M is tiny (<= 32 per pass: 2 x 11|14 GPT-2 tokens, or 2B embedding rows) so the op is bound by streaming W from
// HBM once.  Activations stay fp32-accurate on tensor cores by splitting A = hi + lo (two bf16 terms, ~16 mantissa
// bits) and issuing two warp-level mma.sync per k-step (an mma.sync kernel on purpose: at M <= 32 a tcgen05 tile would be
// 3/4 padding and the op is HBM-bound anyway).  Each quad thread reads 16 contiguous bytes of a weight row (the k index is
// permuted identically for A and B, which leaves the dot product unchanged), so every 32-byte sector fetched is fully used.
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

// ---------------------------------------------------------------- weight-streaming GEMM, M <= 32 rows per pass
// The op is bound by reading W once; everything is organised to keep enough 16-byte weight loads in flight on every SM:
//   * grid (N / 64 column tiles, KS k-slices, M / 32 row passes), 8 warps per CTA: warp w owns output columns [64 bx + 8 w, + 8)
//     over the CTA's k-slice and streams its 8 weight rows straight from global memory, all loads of up to 16 k-steps issued
//     before the first use (no shared memory on the weight path);
//   * the CTA's slice of A (32 rows x the k-slice, fp32) is loaded ONCE, activated, split into hi + lo bf16 and staged in shared
//     memory (rows padded by 64 B: the 16-byte fragment reads of a quarter warp fall into 8 different bank groups) -- round 1
//     re-read and re-split A in every warp for every k-step, ~150 ALU instructions per 512 B of weights;
//   * split-K without a workspace: the KS CTAs of a column tile form a THREAD-BLOCK CLUSTER, leave their 32 x 64 fp32 partials in
//     their own shared memory, and after one cluster barrier CTA r sums columns [64 r / KS, + 64 / KS) of all of them through
//     distributed shared memory in slice order (bit-reproducible), adds bias / activation / residual and writes the result.
// Measured (bench.py --workload c1, GPT-2-medium trunk, 2 x 14 rows): see profiles/README.md.
constexpr int kSmThreads = 256, kSmCols = 64, kSmRows = 32, kSmMaxKR = 1024;      // k-slices longer than kSmMaxKR are walked in chunks

__device__ __forceinline__ float4 ld_dsmem_f32x4(uint32_t cluster_addr) {     // not volatile: the ks loads of a thread are independent
  float4 v;
  asm("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(cluster_addr));
  return v;
}

__global__ void __launch_bounds__(kSmThreads, 2)      // <= 128 registers: two CTAs per SM keep twice the weight loads in flight
gemm_smallm_kernel(const float* __restrict__ A, long long lda, const __nv_bfloat16* __restrict__ W,
                   const float* __restrict__ bias, const float* __restrict__ residual, long long ldr,
                   float* __restrict__ out, long long ldo, int M, int N, int K, int kr, int act_in, int act) {
  // Programmatic dependent launch (common.cuh): this kernel may start while its predecessor is still running.  The WEIGHTS do not
  // depend on the predecessor, so the first 16 k-steps of them are requested before griddepcontrol.wait -- in a chain of tiny
  // dependent GEMMs (the GPT-2 trunk of the prior: 96 per step) the HBM latency of layer i + 1 hides behind layer i.
  pdl_launch_dependents();
  extern __shared__ __align__(16) uint8_t sm_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int ks = (int)gridDim.y, slice = (int)blockIdx.y;         // cluster = the gridDim.y CTAs of one column tile
  const int n0 = blockIdx.x * kSmCols + warp * 8;
  const int m0 = blockIdx.z * kSmRows;
  const int k_lo = slice * kr, k_hi = (k_lo + kr < K) ? k_lo + kr : K;
  const int kc_max = kr < kSmMaxKR ? kr : kSmMaxKR;                // chunk of the slice staged at a time
  const int pitch = kc_max * 2 + 64;                               // bytes per staged row (hi or lo)
  uint8_t* s_hi = sm_raw;
  uint8_t* s_lo = sm_raw + (size_t)kSmRows * pitch;
  const bool col_ok = n0 < N;                                      // N % 8 == 0: a warp's 8 columns exist together
  const __nv_bfloat16* wrow = W + (long long)((col_ok ? n0 : 0) + g) * K + t * 8;
  float acc[2][4];
#pragma unroll
  for (int i = 0; i < 2; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;

  for (int kc0 = k_lo; kc0 < k_hi; kc0 += kc_max) {
    const int kc = (k_hi - kc0 < kc_max) ? k_hi - kc0 : kc_max;    // multiple of 32
    const int nsteps = kc >> 5;
    // ---- the first 16 k-steps of weights go in flight BEFORE A is staged: the HBM latency hides behind the staging work
    uint4 wv[16];
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (col_ok && j < nsteps) wv[j] = __ldg(reinterpret_cast<const uint4*>(wrow + kc0 + j * 32));
    if (kc0 == k_lo) pdl_wait();                                   // A (and bias / residual) come from the predecessor
    // ---- A chunk -> hi / lo bf16 in shared memory (each thread: 8 consecutive k of one row per iteration)
    if (kc0 != k_lo) __syncthreads();                              // previous chunk fully consumed
    for (int i = threadIdx.x; i < kSmRows * (kc >> 3); i += kSmThreads) {
      const int r = i / (kc >> 3), c8 = i - r * (kc >> 3);
      float4 x0 = make_float4(0.f, 0.f, 0.f, 0.f), x1 = x0;
      if (m0 + r < M) {
        const float* ap = A + (long long)(m0 + r) * lda + kc0 + c8 * 8;
        x0 = __ldg(reinterpret_cast<const float4*>(ap));
        x1 = __ldg(reinterpret_cast<const float4*>(ap) + 1);
        if (act_in == IA2P_ACT_SILU) {
          x0.x = silu_f(x0.x); x0.y = silu_f(x0.y); x0.z = silu_f(x0.z); x0.w = silu_f(x0.w);
          x1.x = silu_f(x1.x); x1.y = silu_f(x1.y); x1.z = silu_f(x1.z); x1.w = silu_f(x1.w);
        }
      }
      uint4 h, l;
      split_pair(x0.x, x0.y, h.x, l.x);
      split_pair(x0.z, x0.w, h.y, l.y);
      split_pair(x1.x, x1.y, h.z, l.z);
      split_pair(x1.z, x1.w, h.w, l.w);
      *reinterpret_cast<uint4*>(s_hi + (size_t)r * pitch + c8 * 16) = h;
      *reinterpret_cast<uint4*>(s_lo + (size_t)r * pitch + c8 * 16) = l;
    }
    __syncthreads();
    for (int s0 = 0; s0 < nsteps; s0 += 16) {
      if (s0 > 0) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (col_ok && s0 + j < nsteps) wv[j] = __ldg(reinterpret_cast<const uint4*>(wrow + kc0 + (s0 + j) * 32));
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if (s0 + j < nsteps) {                                     // uniform
          const int ko = (s0 + j) * 64 + t * 16;                   // byte offset of this lane's 8 k inside the staged row
          uint4 ah[4], al[4];                                      // rows g, g + 8, g + 16, g + 24
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            ah[r] = *reinterpret_cast<const uint4*>(s_hi + (size_t)(g + 8 * r) * pitch + ko);
            al[r] = *reinterpret_cast<const uint4*>(s_lo + (size_t)(g + 8 * r) * pitch + ko);
          }
          if (col_ok) {
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
              // k-step A: slots (2t,2t+1) <- pair 0, (2t+8,2t+9) <- pair 1; k-step B: pairs 2, 3 (same permutation on A and W)
              mma_bf16_16816(acc[mt], ah[2 * mt].x, ah[2 * mt + 1].x, ah[2 * mt].y, ah[2 * mt + 1].y, wv[j].x, wv[j].y);
              mma_bf16_16816(acc[mt], al[2 * mt].x, al[2 * mt + 1].x, al[2 * mt].y, al[2 * mt + 1].y, wv[j].x, wv[j].y);
              mma_bf16_16816(acc[mt], ah[2 * mt].z, ah[2 * mt + 1].z, ah[2 * mt].w, ah[2 * mt + 1].w, wv[j].z, wv[j].w);
              mma_bf16_16816(acc[mt], al[2 * mt].z, al[2 * mt + 1].z, al[2 * mt].w, al[2 * mt + 1].w, wv[j].z, wv[j].w);
            }
          }
        }
      }
    }
  }
  // ---- partial tile -> shared memory [32 rows][64 cols] fp32 (over the A staging area; pitch 68 floats keeps rows 16-byte aligned)
  __syncthreads();
  float* part = reinterpret_cast<float*>(sm_raw);
  constexpr int PP = kSmCols + 4;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      const int row = 16 * mt + 8 * hf + g, col = warp * 8 + 2 * t;
      *reinterpret_cast<float2*>(part + row * PP + col) = make_float2(acc[mt][2 * hf], acc[mt][2 * hf + 1]);
    }
  if (ks > 1) cluster_sync_all(); else __syncthreads();
  // ---- CTA `slice` finishes columns [slice * 64 / ks, + 64 / ks), four at a time: sum over the cluster in slice order
  const int cw4 = kSmCols / ks / 4;                                // ks in {1, 2, 4, 8}: 16, 8, 4, 2 column quads
  const uint32_t part_u32 = smem_u32(part);
  for (int i = threadIdx.x; i < kSmRows * cw4; i += kSmThreads) {
    const int row = i / cw4, col = (slice * cw4 + (i - row * cw4)) * 4;
    const int n = blockIdx.x * kSmCols + col, m = m0 + row;
    if (n >= N || m >= M) continue;                                // N % 8 == 0: a quad exists as a whole
    float4 pv[8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
      if (r < ks) pv[r] = (ks == 1) ? *reinterpret_cast<const float4*>(part + row * PP + col)
                                    : ld_dsmem_f32x4(mapa_shared(part_u32 + (uint32_t)(row * PP + col) * 4u, (uint32_t)r));
    float4 v = pv[0];
#pragma unroll
    for (int r = 1; r < 8; ++r)
      if (r < ks) { v.x += pv[r].x; v.y += pv[r].y; v.z += pv[r].z; v.w += pv[r].w; }
    if (bias) {
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + n));
      v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
    }
    v.x = apply_act(v.x, act); v.y = apply_act(v.y, act); v.z = apply_act(v.z, act); v.w = apply_act(v.w, act);
    if (residual) {
      const float* rp = residual + (long long)m * ldr + n;
      v.x += rp[0]; v.y += rp[1]; v.z += rp[2]; v.w += rp[3];
    }
    float* op = out + (long long)m * ldo + n;                      // ldo % 2 == 0 only: two 8-byte stores
    *reinterpret_cast<float2*>(op) = make_float2(v.x, v.y);
    *reinterpret_cast<float2*>(op + 2) = make_float2(v.z, v.w);
  }
  if (ks > 1) cluster_sync_all();                                  // peers may still be reading this CTA's partials
}

// grid (heads, batch); block (32, T): warp i = query i.  Scores: lane j owns key j and computes the whole 64-wide dot product
// itself (independent 16-byte loads, no shuffle chain); softmax over the lanes; output: lane l owns head dims (2l, 2l+1), the
// probabilities arrive by shuffle and the T value loads are independent.  (Round 1 walked the keys serially -- one dependent
// load + 5 shuffles per key: 12.5 us for T = 14 against ~3 us here.)
__global__ void causal_attn_small_kernel(const float* __restrict__ qkv, float* __restrict__ out, int T, int E) {
  pdl_launch_dependents();   // programmatic dependent launch: see common.cuh
  pdl_wait();
  const int i = threadIdx.y, lane = threadIdx.x;
  const int h = blockIdx.x;
  const long long b = blockIdx.y;
  const float* base = qkv + b * T * 3LL * E + h * 64;
  const float4* qp = reinterpret_cast<const float4*>(base + (long long)i * 3 * E);
  const bool live = lane <= i;                                     // causal: key j = lane takes part iff j <= i (i < T <= 32)
  const float4* kp = reinterpret_cast<const float4*>(base + (long long)(live ? lane : 0) * 3 * E + E);
  float s = 0.f;
#pragma unroll
  for (int d = 0; d < 16; ++d) {
    const float4 q4 = __ldg(qp + d), k4 = __ldg(kp + d);
    s += q4.x * k4.x + q4.y * k4.y + q4.z * k4.z + q4.w * k4.w;
  }
  s = live ? s * 0.125f : -INFINITY;
  const float mx = warp_max(s);
  const float p = live ? __expf(s - mx) : 0.f;
  const float sum = warp_sum(p);
  float o0 = 0.f, o1 = 0.f;
  const float* vb = base + 2 * E + 2 * lane;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    if (j <= i) {                                                  // warp-uniform
      const float pj = __shfl_sync(0xffffffffu, p, j);
      const float2 vv = *reinterpret_cast<const float2*>(vb + (long long)j * 3 * E);
      o0 += pj * vv.x;
      o1 += pj * vv.y;
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
  // k-slices: as many (power of two <= 8, each a whole number of 32-wide k-steps and >= 64 long) as it takes to put ~2 CTAs on
  // every SM; the slices of a column tile form one cluster (portable size <= 8)
  const long long tiles = ((N + kSmCols - 1) / kSmCols) * ((M + kSmRows - 1) / kSmRows);
  const int ksteps = (int)(K / 32);
  int ks = 1;
  while (ks < 8 && tiles * ks < 2LL * sm_count() && ksteps % (ks * 2) == 0 && K / (ks * 2) >= 64) ks *= 2;
  const int kr = (int)(K / ks);
  const int kc = kr < kSmMaxKR ? kr : kSmMaxKR;
  const size_t stage = 2 * (size_t)kSmRows * (kc * 2 + 64), part = (size_t)kSmRows * (kSmCols + 4) * sizeof(float);
  const size_t smem = stage > part ? stage : part;
  IA2P_ONCE_PER_DEVICE(IA2P_CUDA(cudaFuncSetAttribute(gemm_smallm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                      (int)(2 * kSmRows * (kSmMaxKR * 2 + 64)))));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)((N + kSmCols - 1) / kSmCols), (unsigned)ks, (unsigned)((M + kSmRows - 1) / kSmRows));
  cfg.blockDim = dim3(kSmThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = (unsigned)ks;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  IA2P_CUDA(cudaLaunchKernelEx(&cfg, gemm_smallm_kernel, A, (long long)lda, static_cast<const __nv_bfloat16*>(W), bias, residual,
                               (long long)ldr, out, (long long)ldo, (int)M, (int)N, (int)K, kr, act_in, act));
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
