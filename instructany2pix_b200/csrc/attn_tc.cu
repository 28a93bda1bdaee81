// tcgen05 / TMEM / TMA flash self-attention for sm_100a, head_dim 64, non-causal (SDPA at attention_processor.py:259-261).
//
// CTA = 128 query rows of one (batch, head); 6 warps: w0 TMA producer, w1 tcgen05.mma issuer (owns TMEM), w2..w5 softmax
// (one query row per thread, so row max / row sum need no shuffles).  Per 128-key block j:
//   MMA1: S(TMEM, 128 cols fp32) = Q(smem) K_j^T(smem)                      [M128 N128 K64]
//   softmax warps: S -> registers (chunk loads software-pipelined against the exp math), p = exp2(s*scale - m_ref) -> bf16
//                  -> TMEM (64 columns, two keys per 32-bit cell: the A operand of MMA2 is read straight from TMEM)
//   MMA2: O(TMEM, 64 cols fp32) += P(TMEM) V_j(smem, MN-major)               [M128 N64 K128]
// O accumulates in TMEM across blocks; the standing reference m_ref (first block: the maximum of its first 32 keys) is only raised
// (and O, l rescaled through tcgen05.ld/st) when a row's block sum shows it is stale by more than 2^20, so the exact-max path is
// rare and the result is exact up to the common factor that cancels in O / l.  MMA1 of block j+1 is issued before MMA2 of block j, and
// two CTAs are co-resident per SM (80 KB smem, 256 TMEM columns each), so tensor work overlaps the softmax math.
// (Round-1 history: P used to go through shared memory -- 16 swizzled 16-byte stores per thread + fence.proxy.async cost
// 450 of the 2 700 cycles a key block took, tools/trace_attn.py; keeping P in TMEM removes that and 32 KB of smem.)
#include <cuda.h>

#include <mutex>

#include "common.cuh"
#include "tensormap.cuh"

namespace ia2p {

struct alignas(64) FaMaps {
  CUtensorMap q, k, v;   // 3-D {heads*64, n_tokens, batch}, box {64, 128, 1}, SWIZZLE_128B
};

constexpr int kFaTile = 128 * 128;                 // bytes: 128 rows x 64 bf16
constexpr int kFaSmemTiles = kFaTile * (1 + 2 + 2);       // Q, K x2, V x2
constexpr int kFaSmemBytes = kFaSmemTiles + 1024;  // barriers live in the alignment slack in front of the tiles
#ifndef IA2P_FA_TRUNC
#define IA2P_FA_TRUNC 1
#endif
#ifndef IA2P_FA_POLY16
#define IA2P_FA_POLY16 5    // of every 16 key PAIRS of the fast path, this many get their exponentials from ex2_poly3_x2 instead of MUFU.EX2
#endif
// What bounds this kernel (measured: tools/sfubench.cu, tools/mmabench.cu "fa pattern", tools/trace_attn.py, profiles/ncu_fa_tc_r02.txt):
//  * tensor pipe: S = Q K^T (4 x M128 N128 K16) and O += P V (8 x M128 N64 K16, A from TMEM) BOTH run at the full rate -- 256 + 256
//    cycles per 128-key block and CTA, 1 036 per block of the two co-resident CTAs (the N = 64 MMAs are not slower per MAC);
//  * softmax: the SM sub-partition dispatches one instruction per cycle, FFMA2 / FADD2 / F2FP occupy two slots, MUFU.EX2 one slot
//    plus 8 cycles of the 4-lane SFU.  Per exponential: scale-subtract 1 + row sum 1 + pack 0.5..1 + (MUFU 1 | polynomial 7) slots,
//    so ~6 cycles per exponential whatever the MUFU / polynomial split between 4/16 and 6/16 (sfubench "softmax mix" rows), i.e.
//    >= 1 540 cycles per block of the CTA pair;
//  * power: under this kernel the chip draws the 1 kW cap (995 W, tools/fa_clocks.py) at ~1.54 GHz, and that is what the variants converge to: one or two
//    threads per row (4 or 8 softmax warps), S in 64-key halves with their own barriers, a staggered start of the co-resident CTAs,
//    truncating or rounding the bf16 pack, 4/16 .. 8/16 polynomial share all land at 840-900 TFLOP/s at 4 096 tokens and 600-660 at
//    1 024 (round 1: 725 / 550; scalar FFMA with 1/4 polynomial: 829 / 588).
// Kept from those experiments: packed fp32 pairs (fewer registers: 126, no spills), P written to TMEM chunk by chunk, the
// first block's reference taken from its first 32 keys (no separate max pass), truncating pack with the 2^-9 bias folded into
// the reference, 5 of 16 pairs on the polynomial.
#ifdef IA2P_TC_TRACE
#define IA2P_TRACE_BUF g_fa_trace
__device__ unsigned long long* g_fa_trace = nullptr;           // debug build: counters of the CTAs with blockIdx.y == z == 0
#endif

__global__ void __launch_bounds__(192, 2)
fa_tc_kernel(const __grid_constant__ FaMaps maps, __nv_bfloat16* __restrict__ out, long long ldo, int n_tokens, float scale_log2) {
  pdl_launch_dependents();   // programmatic dependent launch: see common.cuh
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t tiles = (raw + 128u + 1023u) & ~1023u;       // >= 128 B of barriers in front
  if (tiles + kFaSmemTiles > raw + kFaSmemBytes) __trap();    // dynamic smem base less aligned than assumed
  const uint32_t sQ = tiles, sK = sQ + kFaTile, sV = sK + 2 * kFaTile;
  const uint32_t bars = raw;
  const uint32_t q_full = bars, k_full = bars + 8, v_full = bars + 24, k_empty = bars + 40, v_empty = bars + 56;
  const uint32_t s_full = bars + 72, s_empty = bars + 80, p_full = bars + 88, pv_full = bars + 96, tmem_slot = bars + 104;

  const int warp = warp_idx_uniform(), lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int n_blocks = (n_tokens + 127) >> 7;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.q); tma_prefetch_desc(&maps.k); tma_prefetch_desc(&maps.v);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(k_full + 8 * s, 1); mbar_init(v_full + 8 * s, 1);
      mbar_init(k_empty + 8 * s, 1); mbar_init(v_empty + 8 * s, 1);
    }
    mbar_init(s_full, 1); mbar_init(s_empty, 4); mbar_init(p_full, 4); mbar_init(pv_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
  const uint32_t tS = tmem_base, tO = tmem_base + 128, tP = tmem_base + 192;   // S 128 | O 64 | P 64 (bf16 pairs) columns
  pdl_wait();                // barrier init / TMEM allocation above overlapped the previous kernel's tail

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, kFaTile);
      tma_load_3d(sQ, &maps.q, q_full, h * 64, q0, b);
    }
    for (int j = 0; j < n_blocks; ++j) {
      const int st = j & 1;
      const uint32_t ph = (uint32_t)((j >> 1) & 1);
      mbar_wait(k_empty + 8 * st, ph ^ 1u);
      if (elect_one()) {
        mbar_arrive_expect_tx(k_full + 8 * st, kFaTile);
        tma_load_3d(sK + st * kFaTile, &maps.k, k_full + 8 * st, h * 64, j * 128, b);
      }
      mbar_wait(v_empty + 8 * st, ph ^ 1u);
      if (elect_one()) {
        mbar_arrive_expect_tx(v_full + 8 * st, kFaTile);
        tma_load_3d(sV + st * kFaTile, &maps.v, v_full + 8 * st, h * 64, j * 128, b);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc_qk = umma_idesc_bf16(128, 128, 0);
    constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 64, 1);
    TRACE_DECL(tr_wait_k);
    TRACE_DECL(tr_wait_sempty);
    TRACE_DECL(tr_wait_v);
    TRACE_DECL(tr_wait_p);
    TRACE_T0(tl0);
    auto issue_qk = [&](int j) {
      const int st = j & 1;
      TRACE_T0(ta);
      mbar_wait(k_full + 8 * st, (uint32_t)((j >> 1) & 1));
      TRACE_ADD(tr_wait_k, ta);
      TRACE_T0(tb);
      if (j > 0) mbar_wait(s_empty, (uint32_t)((j - 1) & 1));     // softmax finished reading S_{j-1}
      TRACE_ADD(tr_wait_sempty, tb);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t da = umma_desc_sw128(sQ), db = umma_desc_sw128(sK + st * kFaTile);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tS, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc_qk, (uint32_t)(k != 0));
        umma_commit(k_empty + 8 * st);
        umma_commit(s_full);
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    issue_qk(0);
    for (int j = 0; j < n_blocks; ++j) {
      if (j + 1 < n_blocks) issue_qk(j + 1);                      // overlaps the softmax of block j
      const int st = j & 1;
      TRACE_T0(tc);
      mbar_wait(v_full + 8 * st, (uint32_t)((j >> 1) & 1));
      TRACE_ADD(tr_wait_v, tc);
      TRACE_T0(td);
      mbar_wait(p_full, (uint32_t)(j & 1));                       // P_j in smem; O (TMEM) rescaled if needed
      TRACE_ADD(tr_wait_p, td);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {                          // 8 x 16 keys = 8 TMEM columns of P each
          const uint64_t db = umma_desc_sw128_mn(sV + st * kFaTile + ks * 2048, kFaTile);
          umma_bf16_ts(tO, tP + (uint32_t)(ks * 8), db, idesc_pv, (uint32_t)((j | ks) != 0));
        }
        umma_commit(v_empty + 8 * st);
        umma_commit(pv_full);
      }
      __syncwarp();
    }
#ifdef IA2P_TC_TRACE
    if (blockIdx.y == 0 && blockIdx.z == 0) {
      TRACE_PUT(0, tr_wait_k); TRACE_PUT(1, tr_wait_sempty); TRACE_PUT(2, tr_wait_v); TRACE_PUT(3, tr_wait_p);
      TRACE_PUT(4, clock64() - tl0); TRACE_PUT(9, n_blocks);
    }
#endif
  } else {
    // ------------------------------------------------------------ softmax / epilogue: one query row per thread
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(qd * 32) << 16;
    float m_ref = -INFINITY, l = 0.f;
    TRACE_DECL(tr_wait_s);
    TRACE_DECL(tr_softmax);
    TRACE_DECL(tr_wait_pv);
    TRACE_DECL(tr_pwrite);
    for (int j = 0; j < n_blocks; ++j) {
      TRACE_T0(ts0);
      mbar_wait(s_full, (uint32_t)(j & 1));
      TRACE_ADD(tr_wait_s, ts0);
      TRACE_T0(ts1);
      tc_fence_after();
      const int kv_left = n_tokens - j * 128;                     // valid keys in this block (>= 1)
      const bool ragged = kv_left < 128;                          // only the last block can be ragged
      float lsum = 0.f, alpha = 1.f;
      // P (bf16 pairs) goes to TMEM chunk by chunk, 16 cells per 32 keys, so only one chunk of it is ever live in registers.  The P
      // cells and the O accumulator are free once MMA2 of block j-1 has completed; that MMA was issued right behind MMA1 of this
      // block, so by the first store (a chunk of exponentials after S arrived) the wait is normally over already.
      bool pv_waited = (j == 0);
      auto wait_pv_prev = [&]() {
        if (!pv_waited) {
          TRACE_T0(ts2);
          mbar_wait(pv_full, (uint32_t)((j - 1) & 1));
          tc_fence_after();
          TRACE_ADD(tr_wait_pv, ts2);
          pv_waited = true;
        }
      };
      // FAST PATH: one pass, p = 2^(s*scale - m_ref) against the standing reference max.  bf16 / fp32 share the exponent
      // range, so p may exceed 1 by many orders without harm; the reference is only raised when a row's block sum shows it is
      // stale by > 2^20 (or on a ragged block) -> SLOW PATH below.
      bool slow = ragged;
      if (!slow) {
        // chunk c + 1 is in flight from TMEM while the exponentials of chunk c are computed
        uint32_t va[32], vb[32];
        uint64_t ls0 = f2_pack(0.f, 0.f), ls1 = ls0;
        tmem_ld_32x32(tS + lane_sel, va);
        tmem_ld_wait();
        tmem_ld_32x32(tS + lane_sel + 32u, vb);
        if (j == 0) {
          // First block: ANY reference within ~2^20 of the row maximum will do (it cancels in O / l), so take the maximum of the
          // first 32 keys instead of a separate max pass over all 128; a row whose later keys tower over it fails the sum check
          // below and is redone on the slow path like any other stale reference.
          float mx = __uint_as_float(va[0]);
#pragma unroll
          for (int i = 1; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(va[i]));
          m_ref = mx * scale_log2;
        }
        // IA2P_FA_TRUNC: p is produced a factor (1 + 2^-9) too large (folded into the reference) and TRUNCATED to bf16 by one
        // byte permute per pair instead of a round-to-nearest F2FP (two issue slots): the error is the same +-2^-9 band, centred;
        // the row sum of this block is divided by the factor again below.
        const float nm = IA2P_FA_TRUNC ? 0.00281502f - m_ref : -m_ref;
        const uint64_t sc2 = f2_pack(scale_log2, scale_log2), nm2 = f2_pack(nm, nm);
        auto soft32 = [&](const uint32_t (&v)[32], int c) {
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {                          // pair i = keys 2i, 2i + 1 of this 32-key chunk
            const uint64_t x2 = f2_fma(f2_pack(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), sc2, nm2);
            uint64_t p2;
            if (((i * IA2P_FA_POLY16) & 15) < IA2P_FA_POLY16) {   // IA2P_FA_POLY16 of every 16 pairs on the FMA / ALU pipes
              p2 = ex2_poly3_x2(x2);
            } else {
              float x0, x1;
              f2_unpack(x2, x0, x1);
              p2 = f2_pack(ex2_approx(x0), ex2_approx(x1));
            }
            if (i & 1) ls1 = f2_add(ls1, p2); else ls0 = f2_add(ls0, p2);
            float p0, p1;
            f2_unpack(p2, p0, p1);
            pk[i] = IA2P_FA_TRUNC ? __byte_perm(__float_as_uint(p0), __float_as_uint(p1), 0x7632) : pack_bf16x2(p0, p1);
          }
          wait_pv_prev();
          tmem_st_32x16(tP + lane_sel + (uint32_t)(c * 16), pk);
        };
        soft32(va, 0);
        tmem_ld_wait();
        tmem_ld_32x32(tS + lane_sel + 64u, va);
        soft32(vb, 1);
        tmem_ld_wait();
        tmem_ld_32x32(tS + lane_sel + 96u, vb);
        soft32(va, 2);
        tmem_ld_wait();
        soft32(vb, 3);
        {
          float a0, a1;
          f2_unpack(f2_add(ls0, ls1), a0, a1);
          lsum = a0 + a1;
          if (IA2P_FA_TRUNC) lsum *= 0.998050682f;               // 1 / (1 + 2^-9)
        }
        slow = !(lsum < 1048576.f);                               // also catches inf / nan
      }
      const bool any_slow = __any_sync(0xffffffffu, slow);
      if (any_slow) {
        // SLOW PATH (warp-uniform): exact block row max -> raise the reference, rescale O and l, recompute p (overwrites whatever
        // the fast path stored)
        float mx = -INFINITY;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(tS + lane_sel + (uint32_t)(c * 32), v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (!ragged || c * 32 + i < kv_left) mx = fmaxf(mx, __uint_as_float(v[i]));
        }
        const float m_new = fmaxf(m_ref, mx * scale_log2);
        alpha = (m_ref == -INFINITY) ? 0.f : ex2_approx(m_ref - m_new);
        m_ref = m_new;
        l *= alpha;
        lsum = 0.f;
        tmem_st_wait();                                           // fast-path P stores of this block, if any, are about to be overwritten
        wait_pv_prev();
        if (j > 0) {                                              // rescale O row-wise in TMEM
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            uint32_t v[32];
            tmem_ld_32x32(tO + lane_sel + (uint32_t)(c * 32), v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
            tmem_st_32x32(tO + lane_sel + (uint32_t)(c * 32), v);
          }
        }
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32], pk[16];
          tmem_ld_32x32(tS + lane_sel + (uint32_t)(c * 32), v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float p0 = ex2_approx(fmaf(__uint_as_float(v[i]), scale_log2, -m_ref));
            float p1 = ex2_approx(fmaf(__uint_as_float(v[i + 1]), scale_log2, -m_ref));
            if (ragged) {
              if (c * 32 + i >= kv_left) p0 = 0.f;
              if (c * 32 + i + 1 >= kv_left) p1 = 0.f;
            }
            lsum += p0 + p1;
            pk[i >> 1] = pack_bf16x2(p0, p1);
          }
          tmem_st_32x16(tP + lane_sel + (uint32_t)(c * 16), pk);
        }
      }
      l += lsum;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_empty);                        // S may be overwritten by MMA1 of block j+1
      TRACE_ADD(tr_softmax, ts1);
      TRACE_T0(ts3);
      tmem_st_wait();                                             // this row's P cells (the A operand of MMA2) have landed
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      TRACE_ADD(tr_pwrite, ts3);
    }
#ifdef IA2P_TC_TRACE
    if (blockIdx.y == 0 && blockIdx.z == 0 && warp == 2) {
      TRACE_PUT(5, tr_wait_s); TRACE_PUT(6, tr_softmax); TRACE_PUT(7, tr_wait_pv); TRACE_PUT(8, tr_pwrite);
    }
#endif
    // epilogue: O / l -> bf16 -> global (one 128-byte row per thread)
    mbar_wait(pv_full, (uint32_t)((n_blocks - 1) & 1));
    tc_fence_after();
    const float inv = 1.f / l;
    const bool ok = q0 + row < n_tokens;
    __nv_bfloat16* op = out + ((long long)b * n_tokens + q0 + row) * ldo + h * 64;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      tmem_ld_32x32(tO + lane_sel + (uint32_t)(c * 32), v);
      tmem_ld_wait();
      if (ok) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 o;
          o.x = pack_bf16x2(__uint_as_float(v[8 * i + 0]) * inv, __uint_as_float(v[8 * i + 1]) * inv);
          o.y = pack_bf16x2(__uint_as_float(v[8 * i + 2]) * inv, __uint_as_float(v[8 * i + 3]) * inv);
          o.z = pack_bf16x2(__uint_as_float(v[8 * i + 4]) * inv, __uint_as_float(v[8 * i + 5]) * inv);
          o.w = pack_bf16x2(__uint_as_float(v[8 * i + 6]) * inv, __uint_as_float(v[8 * i + 7]) * inv);
          *reinterpret_cast<uint4*>(op + c * 32 + i * 8) = o;
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

#ifdef IA2P_TC_TRACE
extern "C" int ia2p_debug_set_fa_trace(void* dev_buffer) {
  unsigned long long* p = static_cast<unsigned long long*>(dev_buffer);
  return (int)cudaMemcpyToSymbol(g_fa_trace, &p, sizeof(p));
}
#endif

int launch_fa_tc(const void* q, const void* k, const void* v, int64_t ld, void* out, int64_t ldo, int64_t batch,
                 int64_t n_tokens, int heads, float softmax_scale, cudaStream_t st) {
  FaMaps maps;
  if (int e = make_map_3d_bf16(&maps.q, q, heads * 64, ld, n_tokens, batch, 128, "flash_self_attn")) return e;
  if (int e = make_map_3d_bf16(&maps.k, k, heads * 64, ld, n_tokens, batch, 128, "flash_self_attn")) return e;
  if (int e = make_map_3d_bf16(&maps.v, v, heads * 64, ld, n_tokens, batch, 128, "flash_self_attn")) return e;
  IA2P_ONCE_PER_DEVICE(
      IA2P_CUDA(cudaFuncSetAttribute(fa_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFaSmemBytes));
      IA2P_CUDA(cudaFuncSetAttribute(fa_tc_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)));
  const dim3 grid((unsigned)((n_tokens + 127) / 128), (unsigned)heads, (unsigned)batch);
  launch_pdl(fa_tc_kernel, dim3(grid), dim3(192), kFaSmemBytes, st, maps, static_cast<__nv_bfloat16*>(out), ldo, (int)n_tokens,
                                                softmax_scale * 1.4426950408889634f);
  IA2P_LAUNCH_CHECK();
  return 0;
}

}  // namespace ia2p
