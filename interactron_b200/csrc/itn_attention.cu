// Fused multi-head attention on tcgen05 tensor cores (sm_100a), forward and backward, in the same
// error-compensated tf32x3 arithmetic as itn_gemm_tf32.  The score matrix never leaves the SM:
//   forward   O = softmax(scale * Q K^T + key mask) V            -> O, LSE (one float per row)
//   backward  dQ, dK, dV from Q, K, V, O, dO, LSE                 (scores recomputed per tile)
// See include/interactron_b200.h (itn_attention_fwd / itn_attention_bwd) for the contract and the
// reference call sites (models/gpt.py:43-53, models/detr_models/transformer.py:154-155,219-226).
//
// One CTA owns 128 rows (the 128 TMEM lanes) of one (batch, head) and walks the other sequence in
// blocks of BLK tokens.  Five warps:
//   warps 0-3  "row threads": thread t owns row t.  They load their row operand straight from global
//              memory into TENSOR MEMORY (tcgen05.st; value and tf32 residual), so every MMA of the
//              kernel takes its A operand from TMEM and shared memory only holds the streamed blocks;
//              they split the streamed blocks into value + residual in place, run the softmax /
//              dS arithmetic on the accumulators (tcgen05.ld), write probabilities back to TMEM as
//              the A operand of the next MMA, and keep the output rows in registers.
//   warp 4     controller: one lane issues the TMA loads of the streamed blocks (two stages) and all
//              tcgen05.mma (kind::tf32, 128 x N x 8, A from TMEM, B from shared memory).
// Hand-offs are mbarriers; several CTAs share an SM (TMEM columns permitting) and overlap each
// other's phases.  No atomics: results are bit-reproducible.
//
//   forward        S = Q K^T (B = K block, K-major)          P V      (B = V block, MN-major)
//   backward dQ    S, dP = dO V^T (B = V, K-major)           dS K     (B = K block, MN-major)
//   backward dK/dV S^T = K Q^T, dP^T = V dO^T (B = Q, dO K-major)   P^T dO, dS^T Q (B MN-major)
// The two backward kernels recompute S (and dP) instead of exchanging partial dQ through atomics.
//
// tf32x3: kind::tf32 truncates fp32 operands to 10 mantissa bits.  Every product is issued as
// A*B + A_lo*B + A*B_lo with x_lo = x - trunc_tf32(x) (3 MMAs per k-step), which restores ~fp32
// accuracy.  Accumulation chains are short (3*BLK/8 or 3*HD/8 MMAs; the per-block results are summed in
// fp32 registers), but every tcgen05.mma accumulate rounds toward zero, and even a 24-MMA chain shrinks
// its result by ~4e-7 on average: a systematic scale error on O = P V that the ill-conditioned
// meta-gradients amplify (measured, tools/meta_parity.py: 1.46e-3 with the fused forward against 9.4e-4
// with the compensated GEMM chain).  lo_comp() folds the same first-order compensation as the GEMM's
// splitters (itn_gemm_tf32.cu) into the residual operand that is written to TMEM anyway.
#include <cuda.h>
#include <cuda_runtime.h>

#include <mutex>

#include "itn_common.cuh"
#include "itn_philox.cuh"
#include "itn_ptx.cuh"

namespace itn {
namespace attn {

constexpr int kRows = 128;      // rows per CTA = TMEM lanes
constexpr int kThreads = 160;   // 4 row warps + controller warp
constexpr int kStages = 2;

struct RowView {                // element (b, i, h, d) at p[b*sb + i*ld + h*HD + d]
  float* p;
  long long ld, sb;
};

struct Params {
  RowView q, k, v, o, d_o, dq, dk, dv;
  float* lse;                   // [B*nh, Lq]  log2-sum-exp2 of the scaled scores
  float* delta;                 // [B*nh, Lq]  rowsum(dO * O)
  const unsigned char* kmask;   // [B, Lk], 1 = padded key (may be null)
  int nh, Lq, Lk, tiles;
  float scale, scale_log2;
  DropParams drop;              // train()-mode dropout on the attention probabilities (seed == null: off)
  int dbg;                      // timing experiments (ITN_ATTN_DBG): 1 = K-major descriptors for the MN-major products (wrong results),
                                // 2 = no row arithmetic, 4 = no second-phase MMAs, 8 = no first-phase MMAs, 16 = no residual split
};

// ------------------------------------------------------------------------------- device helpers
__device__ __forceinline__ void wait_bar(uint64_t* bar, uint32_t parity) {
  // a protocol error must not hang the GPU: trap after ~1 s of failed waits
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 24)) __trap();
  }
}

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

#ifndef ITN_ATTN_RZ_EPS
#define ITN_ATTN_RZ_EPS 0.57f   // loss per later round-toward-zero accumulate, in units of 2^-24 ...
#endif
#ifndef ITN_ATTN_RZ_C0
#define ITN_ATTN_RZ_C0 8.5f     // ... plus the chain-length independent shrink (residual operands truncated to 11 bits, lo*lo
#endif                          // dropped): the GEMM's constants (itn_gemm_tf32.cu); fitted 1.2 * 12.5 = 0.57 * 12.5 + 7.9 on 24-MMA chains
#ifndef ITN_ATTN_RZ_COMP
#define ITN_ATTN_RZ_COMP 1
#endif
// Residual (lo) operand of element x at k-index `col` of an NKS-step chain (3 MMAs per step), with the
// round-toward-zero loss of the main product folded in: x * (1 - eps * later accumulates) is what
// survives the chain, so x * eps * (3 * (NKS - ks) - 1) is added back through the residual (<= 1e-6 * x,
// well inside the 2^-11 range of x_lo).  The GEMM's splitters do the same with their own fitted eps.
template <int NKS>
__device__ __forceinline__ float lo_comp(float x, int col) {
#if ITN_ATTN_RZ_COMP
  const float delta = 5.9604645e-8f * (ITN_ATTN_RZ_EPS * (3.0f * static_cast<float>(NKS - (col >> 3)) - 1.0f) + ITN_ATTN_RZ_C0);
  return fmaf(x, delta, tf32_lo(x));
#else
  return tf32_lo(x);
#endif
}

__device__ __forceinline__ float4 tf32_lo4(const float4 v) {
  return make_float4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w));
}

__device__ __forceinline__ constexpr uint32_t tmem_cols(uint32_t need) {
  return need <= 32 ? 32 : need <= 64 ? 64 : need <= 128 ? 128 : need <= 256 ? 256 : 512;
}

// 32 consecutive floats of a row (zeros for a row outside the sequence), times `mul`
__device__ __forceinline__ void load32(uint32_t (&v)[32], const float* src, bool valid, float mul) {
  if (valid) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 x = __ldg(s4 + j);
      v[4 * j + 0] = __float_as_uint(x.x * mul);
      v[4 * j + 1] = __float_as_uint(x.y * mul);
      v[4 * j + 2] = __float_as_uint(x.z * mul);
      v[4 * j + 3] = __float_as_uint(x.w * mul);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = 0u;
  }
}

// this thread's row (HD floats at src) -> TMEM: value at columns t_hi.., tf32 residual at t_lo..
template <int HD>
__device__ __forceinline__ void put_row(uint32_t t_hi, uint32_t t_lo, const float* src, bool valid, float mul) {
#pragma unroll
  for (int c = 0; c < HD / 32; ++c) {
    uint32_t v[32];
    load32(v, src + 32 * c, valid, mul);
    tmem_st_32x32(t_hi + 32 * c, v);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(lo_comp<HD / 8>(__uint_as_float(v[j]), 32 * c + j));
    tmem_st_32x32(t_lo + 32 * c, v);
  }
}

template <int HD>
__device__ __forceinline__ void store_row(float* dst, const float (&acc)[HD], float mul) {
  float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
  for (int j = 0; j < HD / 4; ++j)
    d4[j] = make_float4(acc[4 * j] * mul, acc[4 * j + 1] * mul, acc[4 * j + 2] * mul, acc[4 * j + 3] * mul);
}

// acc[0..HD) += the HD accumulator columns at taddr (this thread's TMEM lane)
template <int HD>
__device__ __forceinline__ void add_cols(float (&acc)[HD], uint32_t taddr) {
#pragma unroll
  for (int c = 0; c < HD / 32; ++c) {
    uint32_t v[32];
    tmem_ld_32x32(taddr + 32 * c, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[32 * c + j] += __uint_as_float(v[j]);
  }
}

// residual tiles of one stage: lo = x - trunc_tf32(x), element-wise on the raw bytes (layout-agnostic,
// so the TMA swizzle is preserved), written hi_bytes further
__device__ __forceinline__ void split_lo(uint8_t* stage, int hi_bytes, int tid) {
  const float4* raw = reinterpret_cast<const float4*>(stage);
  float4* lo = reinterpret_cast<float4*>(stage + hi_bytes);
#pragma unroll 4
  for (int i = tid; i < hi_bytes / 16; i += kRows) lo[i] = tf32_lo4(raw[i]);
}

// K-major streamed tile ([rows][HD], HD contiguous) = HD/32 TMA boxes of [rows x 32 floats], SWIZZLE_128B.
template <int HD>
__device__ __forceinline__ void tma_tile_k(uint8_t* tile, const CUtensorMap* m, uint64_t* bar, int rows,
                                           int row0, int h, int b) {
#pragma unroll
  for (int c = 0; c < HD / 32; ++c) tma_load_4d(tile + c * rows * 128, m, bar, 32 * c, row0, h, b);
}
// descriptor of k-step ks (8 floats of HD) of such a tile used as the B operand [N = rows, K = HD]
__device__ __forceinline__ uint64_t desc_k(uint32_t tile, int rows, int ks) {
  return umma_smem_desc(tile + (ks >> 2) * rows * 128 + (ks & 3) * 32, 16, 1024, kLayoutSW128);
}
// MN-major streamed tile (the same [rows][HD] block used as B operand [N = HD, K = rows]): 4096-byte
// atoms of 32 (HD) x 32 (rows), SWIZZLE_128B_ATOM_32B, atom (kb, jm) at (kb * HD/32 + jm) * 4096.
template <int HD, int BLK>
__device__ __forceinline__ void tma_tile_mn(uint8_t* tile, const CUtensorMap* m, uint64_t* bar, int row0,
                                            int h, int b) {
#pragma unroll
  for (int kb = 0; kb < BLK / 32; ++kb)
#pragma unroll
    for (int jm = 0; jm < HD / 32; ++jm)
      tma_load_4d(tile + (kb * (HD / 32) + jm) * 4096, m, bar, 32 * jm, row0 + 32 * kb, h, b);
}
template <int HD>
__device__ __forceinline__ uint64_t desc_mn(uint32_t tile, int ks) {
  return umma_smem_desc(tile + (ks >> 2) * (HD / 32) * 4096 + (ks & 3) * 1024, 4096, 512, kLayoutSW128Base32);
}

// D[128 x N] (+)= A B with A = (a_hi, a_lo) in TMEM and B = (bd, bd + lo_off) in shared memory
template <int N, bool B_MN>
__device__ __forceinline__ void mma_x3(uint32_t td, uint32_t a_hi, uint32_t a_lo, uint64_t bd, uint64_t lo_off,
                                       bool first) {
  constexpr uint32_t idesc = umma_idesc_tf32(kRows, N, 0, B_MN ? 1 : 0);
  umma_tf32_ts(td, a_hi, bd, idesc, first ? 0u : 1u);
  umma_tf32_ts(td, a_lo, bd, idesc, 1u);
  umma_tf32_ts(td, a_hi, bd + lo_off, idesc, 1u);
}

__device__ __forceinline__ void workers_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

#define ITN_ATTN_PROLOGUE(NBARS)                                                                     \
  extern __shared__ uint8_t smem_raw[];                                                              \
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);                       \
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;                                        \
  const int bh = blockIdx.x / p.tiles, tile = blockIdx.x % p.tiles;                                  \
  const int b = bh / p.nh, h = bh % p.nh;

// =============================================================================================
// forward
// =============================================================================================
template <int HD, int BLK>
struct FwdCfg {
  static constexpr int kTile = BLK * HD * 4;
  static constexpr int kHi = 2 * kTile;                  // K (K-major) | V (MN-major)
  static constexpr int kStage = 2 * kHi;                 // + residual tiles
  static constexpr uint32_t cXH = 0, cXL = HD, cS = 2 * HD, cPL = 2 * HD + BLK, cO = 2 * HD + 2 * BLK;
  static constexpr uint32_t kCols = tmem_cols(3 * HD + 2 * BLK);
  static constexpr int kSmem = kStages * kStage + kStages * BLK * 4 + 16 * 8 + 1024;
};

template <int HD, int BLK, int MINB>
__global__ void __launch_bounds__(kThreads, MINB)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                const __grid_constant__ Params p) {
  using Cfg = FwdCfg<HD, BLK>;
  ITN_ATTN_PROLOGUE()
  float* kbias = reinterpret_cast<float*>(smem + kStages * Cfg::kStage);        // [stage][BLK]: 0 or -inf
  uint64_t* kv_full = reinterpret_cast<uint64_t*>(kbias + kStages * BLK);       // [2] TMA landed
  uint64_t* split_done = kv_full + 2;                                           // [2] residual tiles written
  uint64_t* s_ready = kv_full + 4;                                              // S accumulator complete
  uint64_t* p_ready = kv_full + 5;                                              // P in TMEM (4 warps)
  uint64_t* o_ready = kv_full + 6;                                              // P V complete
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(kv_full + 8);
  const int n_blk = (p.Lk + BLK - 1) / BLK;

  if (threadIdx.x == 128) {
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&split_done[s], 1);
    }
    mbar_init(s_ready, 1);
    mbar_init(p_ready, 4);
    mbar_init(o_ready, 1);
    fence_mbar_init();
  }
  if (warp == 4) tmem_alloc<Cfg::kCols>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tmem_ptr;
  pdl_wait();
  pdl_trigger();

  if (warp == 4) {
    // ------------------------------------------------------------------ controller (converged warp, elect.sync)
    auto issue_tma = [&](int j) {
      const int s = j & 1;
      uint8_t* st = smem + s * Cfg::kStage;
      mbar_expect_tx(&kv_full[s], Cfg::kHi);
      tma_tile_k<HD>(st, &tmK, &kv_full[s], BLK, j * BLK, h, b);
      tma_tile_mn<HD, BLK>(st + Cfg::kTile, &tmV, &kv_full[s], j * BLK, h, b);
    };
    if (elect_one())
      for (int j = 0; j < kStages && j < n_blk; ++j) issue_tma(j);
    __syncwarp();
    constexpr uint64_t kLo = static_cast<uint64_t>(Cfg::kHi >> 4);
    for (int j = 0; j < n_blk; ++j) {
      const int s = j & 1;
      wait_bar(&split_done[s], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t st = smem_u32(smem + s * Cfg::kStage);
      if (elect_one()) {
        if (!(p.dbg & 8))
#pragma unroll
        for (int ks = 0; ks < HD / 8; ++ks)
          mma_x3<BLK, false>(tbase + Cfg::cS, tbase + Cfg::cXH + 8 * ks, tbase + Cfg::cXL + 8 * ks,
                             desc_k(st, BLK, ks), kLo, ks == 0);
        umma_commit(s_ready);
      }
      __syncwarp();
      wait_bar(p_ready, j & 1);
      tc_fence_after();
      if (elect_one()) {
        if (!(p.dbg & 4))
#pragma unroll
        for (int ks = 0; ks < BLK / 8; ++ks)
          mma_x3<HD, true>(tbase + Cfg::cO, tbase + Cfg::cS + 8 * ks, tbase + Cfg::cPL + 8 * ks,
                           desc_mn<HD>(st + Cfg::kTile, ks), kLo, ks == 0);
        umma_commit(o_ready);
      }
      __syncwarp();
      wait_bar(o_ready, j & 1);        // stage s and the S/P columns are free again
      if (j + kStages < n_blk && elect_one()) issue_tma(j + kStages);
      __syncwarp();
    }
  } else {
    // ---------------------------------------------------------------------- row threads
    const int r = threadIdx.x;
    const int i = tile * kRows + r;
    const bool valid = i < p.Lq;
    const uint32_t tw = tbase + (static_cast<uint32_t>(warp * 32) << 16);
    // scores come out of the tensor core in log2 units: q is pre-multiplied by scale * log2(e)
    put_row<HD>(tw + Cfg::cXH, tw + Cfg::cXL, p.q.p + b * p.q.sb + (long long)i * p.q.ld + h * HD, valid,
                p.scale_log2);
    tmem_st_wait();
    auto do_split = [&](int j) {
      const int s = j & 1;
      wait_bar(&kv_full[s], (j >> 1) & 1);
      if (!(p.dbg & 16)) split_lo(smem + s * Cfg::kStage, Cfg::kHi, r);
      if (r < BLK) {
        const int key = j * BLK + r;
        const bool ok = key < p.Lk && !(p.kmask != nullptr && p.kmask[(long long)b * p.Lk + key] != 0);
        kbias[s * BLK + r] = ok ? 0.0f : -INFINITY;
      }
      fence_proxy_async();
      tc_fence_before();
      workers_sync();
      if (r == 0) mbar_arrive(&split_done[s]);
    };
    do_split(0);
    float m_run = -INFINITY, l_run = 0.0f;
    float oacc[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) oacc[d] = 0.0f;
    for (int j = 0; j < n_blk; ++j) {
      const float* kb = kbias + (j & 1) * BLK;
      wait_bar(s_ready, j & 1);
      tc_fence_after();
      float mx = -INFINITY;
      if (!(p.dbg & 2))
#pragma unroll
      for (int c = 0; c < BLK / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(tw + Cfg::cS + 32 * c, v);
        tmem_ld_wait();
#pragma unroll
        for (int t = 0; t < 32; ++t) mx = fmaxf(mx, __uint_as_float(v[t]) + kb[32 * c + t]);
      }
      const float m_new = fmaxf(m_run, mx);
      const float m_use = m_new == -INFINITY ? 0.0f : m_new;     // fully masked so far: keep exp2 finite
      const float alpha = ex2(m_run - m_use);
      float rs = 0.0f;
      if (!(p.dbg & 2))
#pragma unroll
      for (int c = 0; c < BLK / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(tw + Cfg::cS + 32 * c, v);
        tmem_ld_wait();
#pragma unroll
        for (int t = 0; t < 32; ++t) {
          const float pv = ex2(__uint_as_float(v[t]) + kb[32 * c + t] - m_use);
          rs += pv;
          v[t] = __float_as_uint(pv);
        }
        tmem_st_32x32(tw + Cfg::cS + 32 * c, v);                // P over S
#pragma unroll
        for (int t = 0; t < 32; ++t) v[t] = __float_as_uint(lo_comp<BLK / 8>(__uint_as_float(v[t]), 32 * c + t));
        tmem_st_32x32(tw + Cfg::cPL + 32 * c, v);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);
      l_run = l_run * alpha + rs;
      m_run = m_new;
      if (j + 1 < n_blk) do_split(j + 1);                         // overlaps the P V product
#pragma unroll
      for (int d = 0; d < HD; ++d) oacc[d] *= alpha;
      wait_bar(o_ready, j & 1);
      tc_fence_after();
      add_cols<HD>(oacc, tw + Cfg::cO);
    }
    if (valid) {
      const float inv = l_run > 0.0f ? 1.0f / l_run : 0.0f;
      store_row<HD>(p.o.p + b * p.o.sb + (long long)i * p.o.ld + h * HD, oacc, inv);
      p.lse[(long long)bh * p.Lq + i] = l_run > 0.0f ? m_run + log2f(l_run) : INFINITY;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc<Cfg::kCols>(tbase);
  }
}

// =============================================================================================
// backward, dQ (rows = queries, streamed = key blocks); also writes delta = rowsum(dO * O)
// =============================================================================================
template <int HD, int BLK>
struct DqCfg {
  static constexpr int kTile = BLK * HD * 4;
  static constexpr int kHi = 3 * kTile;                  // K (K-major) | V (K-major) | K (MN-major)
  static constexpr int kStage = 2 * kHi;
  static constexpr uint32_t cQH = 0, cQL = HD, cDH = 2 * HD, cDL = 3 * HD, cS = 4 * HD, cDP = 4 * HD + BLK,
                            cDQ = 4 * HD + 2 * BLK;
  static constexpr uint32_t kCols = tmem_cols(5 * HD + 2 * BLK);
  static constexpr int kSmem = kStages * kStage + kStages * BLK * 4 + 16 * 8 + 1024;
};

template <int HD, int BLK, int MINB>
__global__ void __launch_bounds__(kThreads, MINB)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmKk, const __grid_constant__ CUtensorMap tmVk,
                   const __grid_constant__ CUtensorMap tmKm, const __grid_constant__ Params p) {
  using Cfg = DqCfg<HD, BLK>;
  ITN_ATTN_PROLOGUE()
  float* kbias = reinterpret_cast<float*>(smem + kStages * Cfg::kStage);
  uint64_t* kv_full = reinterpret_cast<uint64_t*>(kbias + kStages * BLK);
  uint64_t* split_done = kv_full + 2;
  uint64_t* sdp_ready = kv_full + 4;     // S and dP accumulators complete
  uint64_t* ds_ready = kv_full + 5;      // dS in TMEM (4 warps)
  uint64_t* dq_ready = kv_full + 6;      // dS K complete
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(kv_full + 8);
  const int n_blk = (p.Lk + BLK - 1) / BLK;

  if (threadIdx.x == 128) {
    tma_prefetch_desc(&tmKk);
    tma_prefetch_desc(&tmVk);
    tma_prefetch_desc(&tmKm);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&split_done[s], 1);
    }
    mbar_init(sdp_ready, 1);
    mbar_init(ds_ready, 4);
    mbar_init(dq_ready, 1);
    fence_mbar_init();
  }
  if (warp == 4) tmem_alloc<Cfg::kCols>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tmem_ptr;
  pdl_wait();
  pdl_trigger();

  if (warp == 4) {
    auto issue_tma = [&](int j) {
      const int s = j & 1;
      uint8_t* st = smem + s * Cfg::kStage;
      mbar_expect_tx(&kv_full[s], Cfg::kHi);
      tma_tile_k<HD>(st, &tmKk, &kv_full[s], BLK, j * BLK, h, b);
      tma_tile_k<HD>(st + Cfg::kTile, &tmVk, &kv_full[s], BLK, j * BLK, h, b);
      tma_tile_mn<HD, BLK>(st + 2 * Cfg::kTile, &tmKm, &kv_full[s], j * BLK, h, b);
    };
    if (elect_one())
      for (int j = 0; j < kStages && j < n_blk; ++j) issue_tma(j);
    __syncwarp();
    constexpr uint64_t kLo = static_cast<uint64_t>(Cfg::kHi >> 4);
    for (int j = 0; j < n_blk; ++j) {
      const int s = j & 1;
      wait_bar(&split_done[s], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t st = smem_u32(smem + s * Cfg::kStage);
      if (elect_one()) {
        if (!(p.dbg & 8)) {
#pragma unroll
          for (int ks = 0; ks < HD / 8; ++ks)       // S = Q K^T
            mma_x3<BLK, false>(tbase + Cfg::cS, tbase + Cfg::cQH + 8 * ks, tbase + Cfg::cQL + 8 * ks,
                               desc_k(st, BLK, ks), kLo, ks == 0);
#pragma unroll
          for (int ks = 0; ks < HD / 8; ++ks)       // dP = dO V^T
            mma_x3<BLK, false>(tbase + Cfg::cDP, tbase + Cfg::cDH + 8 * ks, tbase + Cfg::cDL + 8 * ks,
                               desc_k(st + Cfg::kTile, BLK, ks), kLo, ks == 0);
        }
        umma_commit(sdp_ready);
      }
      __syncwarp();
      wait_bar(ds_ready, j & 1);
      tc_fence_after();
      if (elect_one()) {
        if (!(p.dbg & 4))
#pragma unroll
        for (int ks = 0; ks < BLK / 8; ++ks)      // dQ_blk = dS K   (dS value over dP, residual over S)
          mma_x3<HD, true>(tbase + Cfg::cDQ, tbase + Cfg::cDP + 8 * ks, tbase + Cfg::cS + 8 * ks,
                           desc_mn<HD>(st + 2 * Cfg::kTile, ks), kLo, ks == 0);
        umma_commit(dq_ready);
      }
      __syncwarp();
      wait_bar(dq_ready, j & 1);
      if (j + kStages < n_blk && elect_one()) issue_tma(j + kStages);
      __syncwarp();
    }
  } else {
    const int r = threadIdx.x;
    const int i = tile * kRows + r;
    const bool valid = i < p.Lq;
    const uint32_t tw = tbase + (static_cast<uint32_t>(warp * 32) << 16);
    const float* qrow = p.q.p + b * p.q.sb + (long long)i * p.q.ld + h * HD;
    const float* dorow = p.d_o.p + b * p.d_o.sb + (long long)i * p.d_o.ld + h * HD;
    put_row<HD>(tw + Cfg::cQH, tw + Cfg::cQL, qrow, valid, p.scale_log2);
    put_row<HD>(tw + Cfg::cDH, tw + Cfg::cDL, dorow, valid, 1.0f);
    float delta = 0.0f, lse2 = INFINITY;
    if (valid) {
      const float4* a4 = reinterpret_cast<const float4*>(dorow);
      const float4* o4 = reinterpret_cast<const float4*>(p.o.p + b * p.o.sb + (long long)i * p.o.ld + h * HD);
#pragma unroll
      for (int j = 0; j < HD / 4; ++j) {
        const float4 x = __ldg(a4 + j), y = __ldg(o4 + j);
        delta = fmaf(x.x, y.x, delta);
        delta = fmaf(x.y, y.y, delta);
        delta = fmaf(x.z, y.z, delta);
        delta = fmaf(x.w, y.w, delta);
      }
      lse2 = p.lse[(long long)bh * p.Lq + i];
      p.delta[(long long)bh * p.Lq + i] = delta;
    }
    tmem_st_wait();
    auto do_split = [&](int j) {
      const int s = j & 1;
      wait_bar(&kv_full[s], (j >> 1) & 1);
      if (!(p.dbg & 16)) split_lo(smem + s * Cfg::kStage, Cfg::kHi, r);
      if (r < BLK) {
        const int key = j * BLK + r;
        const bool ok = key < p.Lk && !(p.kmask != nullptr && p.kmask[(long long)b * p.Lk + key] != 0);
        kbias[s * BLK + r] = ok ? 0.0f : -INFINITY;
      }
      fence_proxy_async();
      tc_fence_before();
      workers_sync();
      if (r == 0) mbar_arrive(&split_done[s]);
    };
    do_split(0);
    float dqacc[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) dqacc[d] = 0.0f;
    for (int j = 0; j < n_blk; ++j) {
      const float* kb = kbias + (j & 1) * BLK;
      wait_bar(sdp_ready, j & 1);
      tc_fence_after();
      if (!(p.dbg & 2))
#pragma unroll
      for (int c = 0; c < BLK / 32; ++c) {
        uint32_t sv[32], dp[32];
        tmem_ld_32x32(tw + Cfg::cS + 32 * c, sv);
        tmem_ld_32x32(tw + Cfg::cDP + 32 * c, dp);
        tmem_ld_wait();
#pragma unroll
        for (int t = 0; t < 32; ++t) {
          const float pv = ex2(__uint_as_float(sv[t]) + kb[32 * c + t] - lse2);
          const float ds = pv * (__uint_as_float(dp[t]) - delta) * p.scale;
          dp[t] = __float_as_uint(ds);
          sv[t] = __float_as_uint(lo_comp<BLK / 8>(ds, 32 * c + t));
        }
        tmem_st_32x32(tw + Cfg::cDP + 32 * c, dp);
        tmem_st_32x32(tw + Cfg::cS + 32 * c, sv);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(ds_ready);
      if (j + 1 < n_blk) do_split(j + 1);
      wait_bar(dq_ready, j & 1);
      tc_fence_after();
      add_cols<HD>(dqacc, tw + Cfg::cDQ);
    }
    if (valid) store_row<HD>(p.dq.p + b * p.dq.sb + (long long)i * p.dq.ld + h * HD, dqacc, 1.0f);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc<Cfg::kCols>(tbase);
  }
}

// =============================================================================================
// backward, dK / dV (rows = keys, streamed = query blocks)
// =============================================================================================
template <int HD, int BLK>
struct DkvCfg {
  static constexpr int kTile = BLK * HD * 4;
  static constexpr int kHi = 4 * kTile;                  // Q (K-major) | dO (K-major) | Q (MN-major) | dO (MN-major)
  static constexpr int kStage = 2 * kHi;
  static constexpr uint32_t cKH = 0, cKL = HD, cVH = 2 * HD, cVL = 3 * HD, cST = 4 * HD, cPL = 4 * HD + BLK,
                            cDPT = 4 * HD + 2 * BLK, cDSL = 4 * HD + 3 * BLK, cDV = 4 * HD + 4 * BLK,
                            cDK = 5 * HD + 4 * BLK;
  static constexpr uint32_t kCols = tmem_cols(6 * HD + 4 * BLK);
  static_assert(6 * HD + 4 * BLK <= 512, "tensor memory has 512 columns");
  static constexpr int kSmem = kStages * kStage + 2 * kStages * BLK * 4 + 16 * 8 + 1024;
};

template <int HD, int BLK, int MINB, bool DROP>
__global__ void __launch_bounds__(kThreads, MINB)
attn_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tmQk, const __grid_constant__ CUtensorMap tmDOk,
                    const __grid_constant__ CUtensorMap tmQm, const __grid_constant__ CUtensorMap tmDOm,
                    const __grid_constant__ Params p) {
  using Cfg = DkvCfg<HD, BLK>;
  ITN_ATTN_PROLOGUE()
  float* lse_s = reinterpret_cast<float*>(smem + kStages * Cfg::kStage);     // [stage][BLK]
  float* dl_s = lse_s + kStages * BLK;                                       // [stage][BLK]
  uint64_t* q_full = reinterpret_cast<uint64_t*>(dl_s + kStages * BLK);
  uint64_t* split_done = q_full + 2;
  uint64_t* st_ready = q_full + 4;       // S^T and dP^T accumulators complete
  uint64_t* p_ready = q_full + 5;        // P^T, dS^T in TMEM (4 warps)
  uint64_t* dkv_ready = q_full + 6;      // P^T dO and dS^T Q complete
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(q_full + 8);
  const int n_blk = (p.Lq + BLK - 1) / BLK;

  if (threadIdx.x == 128) {
    tma_prefetch_desc(&tmQk);
    tma_prefetch_desc(&tmDOk);
    tma_prefetch_desc(&tmQm);
    tma_prefetch_desc(&tmDOm);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&q_full[s], 1);
      mbar_init(&split_done[s], 1);
    }
    mbar_init(st_ready, 1);
    mbar_init(p_ready, 4);
    mbar_init(dkv_ready, 1);
    fence_mbar_init();
  }
  if (warp == 4) tmem_alloc<Cfg::kCols>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tmem_ptr;
  pdl_wait();
  pdl_trigger();

  if (warp == 4) {
    auto issue_tma = [&](int j) {
      const int s = j & 1;
      uint8_t* st = smem + s * Cfg::kStage;
      mbar_expect_tx(&q_full[s], Cfg::kHi);
      tma_tile_k<HD>(st, &tmQk, &q_full[s], BLK, j * BLK, h, b);
      tma_tile_k<HD>(st + Cfg::kTile, &tmDOk, &q_full[s], BLK, j * BLK, h, b);
      tma_tile_mn<HD, BLK>(st + 2 * Cfg::kTile, &tmQm, &q_full[s], j * BLK, h, b);
      tma_tile_mn<HD, BLK>(st + 3 * Cfg::kTile, &tmDOm, &q_full[s], j * BLK, h, b);
    };
    if (elect_one())
      for (int j = 0; j < kStages && j < n_blk; ++j) issue_tma(j);
    __syncwarp();
    constexpr uint64_t kLo = static_cast<uint64_t>(Cfg::kHi >> 4);
    for (int j = 0; j < n_blk; ++j) {
      const int s = j & 1;
      wait_bar(&split_done[s], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t st = smem_u32(smem + s * Cfg::kStage);
      if (elect_one()) {
        if (!(p.dbg & 8)) {
#pragma unroll
          for (int ks = 0; ks < HD / 8; ++ks)       // S^T = K Q^T
            mma_x3<BLK, false>(tbase + Cfg::cST, tbase + Cfg::cKH + 8 * ks, tbase + Cfg::cKL + 8 * ks,
                               desc_k(st, BLK, ks), kLo, ks == 0);
#pragma unroll
          for (int ks = 0; ks < HD / 8; ++ks)       // dP^T = V dO^T
            mma_x3<BLK, false>(tbase + Cfg::cDPT, tbase + Cfg::cVH + 8 * ks, tbase + Cfg::cVL + 8 * ks,
                               desc_k(st + Cfg::kTile, BLK, ks), kLo, ks == 0);
        }
        umma_commit(st_ready);
      }
      __syncwarp();
      wait_bar(p_ready, j & 1);
      tc_fence_after();
      if (elect_one()) {
        if (!(p.dbg & 4)) {
#pragma unroll
          for (int ks = 0; ks < BLK / 8; ++ks)      // dV_blk = P^T dO
            mma_x3<HD, true>(tbase + Cfg::cDV, tbase + Cfg::cST + 8 * ks, tbase + Cfg::cPL + 8 * ks,
                             desc_mn<HD>(st + 3 * Cfg::kTile, ks), kLo, ks == 0);
#pragma unroll
          for (int ks = 0; ks < BLK / 8; ++ks)      // dK_blk = dS^T Q
            mma_x3<HD, true>(tbase + Cfg::cDK, tbase + Cfg::cDPT + 8 * ks, tbase + Cfg::cDSL + 8 * ks,
                             desc_mn<HD>(st + 2 * Cfg::kTile, ks), kLo, ks == 0);
        }
        umma_commit(dkv_ready);
      }
      __syncwarp();
      wait_bar(dkv_ready, j & 1);
      if (j + kStages < n_blk && elect_one()) issue_tma(j + kStages);
      __syncwarp();
    }
  } else {
    const int r = threadIdx.x;
    const int key = tile * kRows + r;
    const bool inside = key < p.Lk;
    const bool kvalid = inside && !(p.kmask != nullptr && p.kmask[(long long)b * p.Lk + key] != 0);
    const uint32_t tw = tbase + (static_cast<uint32_t>(warp * 32) << 16);
    put_row<HD>(tw + Cfg::cKH, tw + Cfg::cKL, p.k.p + b * p.k.sb + (long long)key * p.k.ld + h * HD, inside,
                p.scale_log2);
    put_row<HD>(tw + Cfg::cVH, tw + Cfg::cVL, p.v.p + b * p.v.sb + (long long)key * p.v.ld + h * HD, inside, 1.0f);
    tmem_st_wait();
    auto do_split = [&](int j) {
      const int s = j & 1;
      wait_bar(&q_full[s], (j >> 1) & 1);
      if (!(p.dbg & 16)) split_lo(smem + s * Cfg::kStage, Cfg::kHi, r);
      if (r < BLK) {
        const int qi = j * BLK + r;
        const bool ok = qi < p.Lq;
        lse_s[s * BLK + r] = ok ? p.lse[(long long)bh * p.Lq + qi] : INFINITY;
        dl_s[s * BLK + r] = ok ? p.delta[(long long)bh * p.Lq + qi] : 0.0f;
      }
      fence_proxy_async();
      tc_fence_before();
      workers_sync();
      if (r == 0) mbar_arrive(&split_done[s]);
    };
    do_split(0);
    float dvacc[HD], dkacc[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) dvacc[d] = dkacc[d] = 0.0f;
    constexpr bool drop = DROP;
    const unsigned long long dseed = DROP ? *p.drop.seed : 0ull;
    const unsigned long long qrow0 = (unsigned long long)bh * p.Lq;
    for (int j = 0; j < n_blk; ++j) {
      const float* ls = lse_s + (j & 1) * BLK;
      const float* dl = dl_s + (j & 1) * BLK;
      wait_bar(st_ready, j & 1);
      tc_fence_after();
      if (!(p.dbg & 2))
#pragma unroll
      for (int c = 0; c < BLK / 32; ++c) {
        uint32_t sv[32], dp[32];
        tmem_ld_32x32(tw + Cfg::cST + 32 * c, sv);
        tmem_ld_32x32(tw + Cfg::cDPT + 32 * c, dp);
        tmem_ld_wait();
#pragma unroll
        for (int t = 0; t < 32; ++t) {
          const float pv = kvalid ? ex2(__uint_as_float(sv[t]) - ls[32 * c + t]) : 0.0f;
          float dpv = __uint_as_float(dp[t]), pd = pv;
          if (drop) {         // rows are keys here: the mask word of (query row, this key) is word key % 4
            const uint4 w4 = dropout_words(dseed, p.drop.site, qrow0 + (unsigned)(j * BLK + 32 * c + t), (unsigned)key >> 2);
            const float kf = word_of(w4, key & 3) >= p.drop.thr ? p.drop.inv_keep : 0.0f;
            pd = pv * kf;
            dpv *= kf;
          }
          const float ds = pv * (dpv - dl[32 * c + t]) * p.scale;
          sv[t] = __float_as_uint(pd);
          dp[t] = __float_as_uint(ds);
        }
        tmem_st_32x32(tw + Cfg::cST + 32 * c, sv);              // P^T over S^T
        tmem_st_32x32(tw + Cfg::cDPT + 32 * c, dp);             // dS^T over dP^T
#pragma unroll
        for (int t = 0; t < 32; ++t) {
          sv[t] = __float_as_uint(lo_comp<BLK / 8>(__uint_as_float(sv[t]), 32 * c + t));
          dp[t] = __float_as_uint(lo_comp<BLK / 8>(__uint_as_float(dp[t]), 32 * c + t));
        }
        tmem_st_32x32(tw + Cfg::cPL + 32 * c, sv);
        tmem_st_32x32(tw + Cfg::cDSL + 32 * c, dp);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);
      if (j + 1 < n_blk) do_split(j + 1);
      wait_bar(dkv_ready, j & 1);
      tc_fence_after();
      add_cols<HD>(dvacc, tw + Cfg::cDV);
      add_cols<HD>(dkacc, tw + Cfg::cDK);
    }
    if (inside) {
      store_row<HD>(p.dv.p + b * p.dv.sb + (long long)key * p.dv.ld + h * HD, dvacc, 1.0f);
      store_row<HD>(p.dk.p + b * p.dk.sb + (long long)key * p.dk.ld + h * HD, dkacc, 1.0f);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc<Cfg::kCols>(tbase);
  }
}

// =============================================================================================
// Software-pipelined variant (all three passes), 8 warps:
//   warps 0-3  row threads (arithmetic only)      warp 4  TMA producer (NS stages ahead)
//   warp 5     TMEM owner + MMA issuer             warps 6-7  residual splitters (+ key-bias / LSE vectors)
// The first-phase accumulators (S, dP, ...) are DOUBLE-BUFFERED in tensor memory: the issuer puts the
// products of block j+1 on the tensor pipe before the row threads have finished block j, and the
// second-phase results (P V, dS K, ...) are read back one block late (double-buffered as well), so
// neither side waits for the other in steady state.  tcgen05.mma of one thread execute in issue order,
// which is what makes S(j+2) overwriting the probabilities of block j safe: the product that reads
// them was issued first.
//   MODE 0 forward           rows = queries   streamed: K (K-major), V (MN-major)
//   MODE 1 backward dQ       rows = queries   streamed: K, V (K-major), K (MN-major)
//   MODE 2 backward dK/dV    rows = keys      streamed: Q, dO (K-major), Q, dO (MN-major)
// =============================================================================================
enum { M_FWD = 0, M_DQ = 1, M_DKV = 2 };
constexpr int kPipeThreads = 256;

template <int MODE, int HD, int BLK>
struct PipeCfg {
  static constexpr int kNK = MODE == M_FWD ? 1 : 2;        // K-major streamed tiles
  static constexpr int kNM = MODE == M_DKV ? 2 : 1;        // MN-major streamed tiles
  static constexpr int kNX = MODE == M_FWD ? 1 : 2;        // row operands resident in TMEM (value + residual)
  static constexpr int kNO = MODE == M_DKV ? 2 : 1;        // outputs
  static constexpr int kSetCols = (MODE == M_DKV ? 4 : 2) * BLK;   // one first-phase buffer set
  static constexpr int kTile = BLK * HD * 4;
  static constexpr int kHi = (kNK + kNM) * kTile;
  static constexpr int kStage = 2 * kHi;
  static constexpr int kMaxNS = (227 * 1024 - 4096) / kStage;
  static constexpr int kNS = kMaxNS > 4 ? 4 : kMaxNS;
  static_assert(kNS >= 2, "need at least two stages");
  static constexpr uint32_t cX = 0;                        // operand x: value at cX + 2*HD*x, residual HD further
  static constexpr uint32_t cSet = 2 * HD * kNX;           // buffer set i at cSet + i*kSetCols
  static constexpr uint32_t cOut = cSet + 2 * kSetCols;    // output buffer i at cOut + i*kNO*HD
  static constexpr uint32_t kNeed = cOut + 2 * kNO * HD;
  static_assert(kNeed <= 512, "tensor memory has 512 columns");
  static constexpr uint32_t kCols = tmem_cols(kNeed);
  static constexpr int kVecBytes = kNS * BLK * 4 * 2;      // two float vectors per stage
  static constexpr int kBars = 3 * kNS + 9;
  static constexpr int kSmem = kNS * kStage + kVecBytes + kBars * 8 + 16 + 1024;
  static_assert(kSmem <= 227 * 1024, "exceeds shared memory per CTA");
};

template <int MODE, int HD, int BLK, int MINB, bool DROP>
__global__ void __launch_bounds__(kPipeThreads, MINB)
attn_pipe_kernel(const __grid_constant__ CUtensorMap tm0, const __grid_constant__ CUtensorMap tm1,
                 const __grid_constant__ CUtensorMap tm2, const __grid_constant__ CUtensorMap tm3,
                 const __grid_constant__ Params p) {
  using Cfg = PipeCfg<MODE, HD, BLK>;
  constexpr int NS = Cfg::kNS;
  ITN_ATTN_PROLOGUE()
  float* vec0 = reinterpret_cast<float*>(smem + NS * Cfg::kStage);   // [NS][BLK]  key bias (fwd, dQ) / LSE (dK,dV)
  float* vec1 = vec0 + NS * BLK;                                     // [NS][BLK]  delta (dK,dV)
  uint64_t* full = reinterpret_cast<uint64_t*>(vec1 + NS * BLK);     // [NS] TMA landed
  uint64_t* split = full + NS;                                       // [NS] residual tiles + vectors written (2 warps)
  uint64_t* free_ = split + NS;                                      // [NS] second-phase MMAs of the block retired
  uint64_t* p1_ready = free_ + NS;                                   // [2] first-phase accumulators complete
  uint64_t* a2_ready = p1_ready + 2;                                 // [2] second-phase A operands in TMEM (4 warps)
  uint64_t* out_ready = a2_ready + 2;                                // [2] second-phase products complete
  uint64_t* out_free = out_ready + 2;                                // [2] row threads have read them (4 warps)
  uint64_t* x_ready = out_free + 2;                                  // row operands in TMEM (4 warps)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(x_ready + 1);
  const int n_blk = ((MODE == M_DKV ? p.Lq : p.Lk) + BLK - 1) / BLK;

  if (threadIdx.x == 128) {
    tma_prefetch_desc(&tm0);
    tma_prefetch_desc(&tm1);
    if (MODE != M_FWD) tma_prefetch_desc(&tm2);
    if (MODE == M_DKV) tma_prefetch_desc(&tm3);
    for (int s = 0; s < NS; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&split[s], 2);
      mbar_init(&free_[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&p1_ready[i], 1);
      mbar_init(&a2_ready[i], 4);
      mbar_init(&out_ready[i], 1);
      mbar_init(&out_free[i], 4);
    }
    mbar_init(x_ready, 4);
    fence_mbar_init();
  }
  if (warp == 5) tmem_alloc<Cfg::kCols>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tmem_ptr;
  pdl_wait();
  pdl_trigger();

  if (warp == 4) {
    // ------------------------------------------------------------------ TMA producer (converged warp)
    for (int j = 0; j < n_blk; ++j) {
      const int s = j % NS, k = j / NS;
      if (k > 0) wait_bar(&free_[s], (k - 1) & 1);
      if (elect_one()) {
        uint8_t* st = smem + s * Cfg::kStage;
        mbar_expect_tx(&full[s], Cfg::kHi);
        tma_tile_k<HD>(st, &tm0, &full[s], BLK, j * BLK, h, b);
        if (MODE == M_FWD) {
          tma_tile_mn<HD, BLK>(st + Cfg::kTile, &tm1, &full[s], j * BLK, h, b);
        } else {
          tma_tile_k<HD>(st + Cfg::kTile, &tm1, &full[s], BLK, j * BLK, h, b);
          tma_tile_mn<HD, BLK>(st + 2 * Cfg::kTile, &tm2, &full[s], j * BLK, h, b);
          if (MODE == M_DKV) tma_tile_mn<HD, BLK>(st + 3 * Cfg::kTile, &tm3, &full[s], j * BLK, h, b);
        }
      }
      __syncwarp();
    }
  } else if (warp == 5) {
    // ------------------------------------------------------------------ MMA issuer (converged warp, elect.sync)
    constexpr uint64_t kLo = static_cast<uint64_t>(Cfg::kHi >> 4);
    auto phase2 = [&](int i) {
      const int set = i & 1;
      wait_bar(&a2_ready[set], (i >> 1) & 1);
      if (i >= 2) wait_bar(&out_free[set], ((i >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t st = smem_u32(smem + (i % NS) * Cfg::kStage);
      const uint32_t ts = tbase + Cfg::cSet + set * Cfg::kSetCols;
      const uint32_t to = tbase + Cfg::cOut + set * Cfg::kNO * HD;
      if (elect_one()) {
        if (!(p.dbg & 4)) {
          if (MODE == M_FWD) {
#pragma unroll
            for (int ks = 0; ks < BLK / 8; ++ks)      // O_blk = P V
              mma_x3<HD, true>(to, ts + 8 * ks, ts + BLK + 8 * ks, desc_mn<HD>(st + Cfg::kTile, ks), kLo, ks == 0);
          } else if (MODE == M_DQ) {
#pragma unroll
            for (int ks = 0; ks < BLK / 8; ++ks)      // dQ_blk = dS K   (dS value over dP, residual over S)
              mma_x3<HD, true>(to, ts + BLK + 8 * ks, ts + 8 * ks, desc_mn<HD>(st + 2 * Cfg::kTile, ks), kLo, ks == 0);
          } else {
#pragma unroll
            for (int ks = 0; ks < BLK / 8; ++ks)      // dV_blk = P^T dO
              mma_x3<HD, true>(to, ts + 8 * ks, ts + BLK + 8 * ks, desc_mn<HD>(st + 3 * Cfg::kTile, ks), kLo, ks == 0);
#pragma unroll
            for (int ks = 0; ks < BLK / 8; ++ks)      // dK_blk = dS^T Q
              mma_x3<HD, true>(to + HD, ts + 2 * BLK + 8 * ks, ts + 3 * BLK + 8 * ks,
                               desc_mn<HD>(st + 2 * Cfg::kTile, ks), kLo, ks == 0);
          }
        }
        umma_commit(&out_ready[set]);
        umma_commit(&free_[i % NS]);
      }
      __syncwarp();
    };
    wait_bar(x_ready, 0);
    for (int j = 0; j < n_blk; ++j) {
      const int s = j % NS;
      wait_bar(&split[s], (j / NS) & 1);
      tc_fence_after();
      const uint32_t st = smem_u32(smem + s * Cfg::kStage);
      const uint32_t ts = tbase + Cfg::cSet + (j & 1) * Cfg::kSetCols;
      const uint32_t x0 = tbase + Cfg::cX, x1 = tbase + Cfg::cX + 2 * HD;
      if (elect_one()) {
        if (!(p.dbg & 8)) {
#pragma unroll
          for (int ks = 0; ks < HD / 8; ++ks)         // S = Q K^T   |   S^T = K Q^T
            mma_x3<BLK, false>(ts, x0 + 8 * ks, x0 + HD + 8 * ks, desc_k(st, BLK, ks), kLo, ks == 0);
          if (MODE != M_FWD) {
            const uint32_t td = ts + (MODE == M_DQ ? BLK : 2 * BLK);
#pragma unroll
            for (int ks = 0; ks < HD / 8; ++ks)       // dP = dO V^T  |  dP^T = V dO^T
              mma_x3<BLK, false>(td, x1 + 8 * ks, x1 + HD + 8 * ks, desc_k(st + Cfg::kTile, BLK, ks), kLo, ks == 0);
          }
        }
        umma_commit(&p1_ready[j & 1]);
      }
      __syncwarp();
      if (j >= 1) phase2(j - 1);
    }
    phase2(n_blk - 1);
  } else if (warp >= 6) {
    // ------------------------------------------------------------------ residual splitters
    const int tid = threadIdx.x - 192;      // 0..63
    for (int j = 0; j < n_blk; ++j) {
      const int s = j % NS;
      wait_bar(&full[s], (j / NS) & 1);
      if (!(p.dbg & 16)) {
        const float4* raw = reinterpret_cast<const float4*>(smem + s * Cfg::kStage);
        float4* lo = reinterpret_cast<float4*>(smem + s * Cfg::kStage + Cfg::kHi);
#pragma unroll 4
        for (int i = tid; i < Cfg::kHi / 16; i += 64) lo[i] = tf32_lo4(raw[i]);
      }
      for (int t = tid; t < BLK; t += 64) {
        const int idx = j * BLK + t;
        if (MODE == M_DKV) {
          const bool ok = idx < p.Lq;
          vec0[s * BLK + t] = ok ? p.lse[(long long)bh * p.Lq + idx] : INFINITY;
          vec1[s * BLK + t] = ok ? p.delta[(long long)bh * p.Lq + idx] : 0.0f;
        } else {
          const bool ok = idx < p.Lk && !(p.kmask != nullptr && p.kmask[(long long)b * p.Lk + idx] != 0);
          vec0[s * BLK + t] = ok ? 0.0f : -INFINITY;
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&split[s]);
    }
  } else {
    // ------------------------------------------------------------------ row threads
    const int r = threadIdx.x;
    const int i = tile * kRows + r;
    const uint32_t tw = tbase + (static_cast<uint32_t>(warp * 32) << 16);
    const uint32_t x0 = tw + Cfg::cX, x1 = tw + Cfg::cX + 2 * HD;
    bool valid;
    float lse2 = INFINITY, delta = 0.0f;
    if (MODE == M_DKV) {
      const bool inside = i < p.Lk;
      valid = inside && !(p.kmask != nullptr && p.kmask[(long long)b * p.Lk + i] != 0);
      put_row<HD>(x0, x0 + HD, p.k.p + b * p.k.sb + (long long)i * p.k.ld + h * HD, inside, p.scale_log2);
      put_row<HD>(x1, x1 + HD, p.v.p + b * p.v.sb + (long long)i * p.v.ld + h * HD, inside, 1.0f);
    } else {
      valid = i < p.Lq;
      put_row<HD>(x0, x0 + HD, p.q.p + b * p.q.sb + (long long)i * p.q.ld + h * HD, valid, p.scale_log2);
      if (MODE == M_DQ) {
        const float* dorow = p.d_o.p + b * p.d_o.sb + (long long)i * p.d_o.ld + h * HD;
        put_row<HD>(x1, x1 + HD, dorow, valid, 1.0f);
        if (valid) {
          const float4* a4 = reinterpret_cast<const float4*>(dorow);
          const float4* o4 = reinterpret_cast<const float4*>(p.o.p + b * p.o.sb + (long long)i * p.o.ld + h * HD);
#pragma unroll
          for (int t = 0; t < HD / 4; ++t) {
            const float4 x = __ldg(a4 + t), y = __ldg(o4 + t);
            delta = fmaf(x.x, y.x, delta);
            delta = fmaf(x.y, y.y, delta);
            delta = fmaf(x.z, y.z, delta);
            delta = fmaf(x.w, y.w, delta);
          }
          lse2 = p.lse[(long long)bh * p.Lq + i];
          p.delta[(long long)bh * p.Lq + i] = delta;
        }
      }
    }
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(x_ready);

    float acc0[HD];
    float acc1[MODE == M_DKV ? HD : 1];
#pragma unroll
    for (int d = 0; d < HD; ++d) acc0[d] = 0.0f;
    if (MODE == M_DKV) {
#pragma unroll
      for (int d = 0; d < (MODE == M_DKV ? HD : 1); ++d) acc1[d] = 0.0f;
    }
    float m_run = -INFINITY, l_run = 0.0f, alpha_pend = 1.0f;
    // train()-mode dropout of the probabilities: keep(seed, site, query row, key) regenerated per element
    // (compile-time switch: the eval() kernels carry none of it)
    constexpr bool drop = DROP;
    const unsigned long long dseed = DROP ? *p.drop.seed : 0ull;
    const unsigned long long qrow0 = (unsigned long long)bh * p.Lq;     // + query index = mask row

    auto consume = [&](int blk) {            // read back the second-phase products of block `blk`
      const int set = blk & 1;
      wait_bar(&out_ready[set], (blk >> 1) & 1);
      tc_fence_after();
      const uint32_t to = tw + Cfg::cOut + set * Cfg::kNO * HD;
      if (MODE == M_FWD) {
#pragma unroll
        for (int d = 0; d < HD; ++d) acc0[d] *= alpha_pend;
      }
      add_cols<HD>(acc0, to);
      if (MODE == M_DKV) add_cols<(MODE == M_DKV ? HD : 1)>(acc1, to + HD);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&out_free[set]);
    };

    for (int j = 0; j < n_blk; ++j) {
      const int s = j % NS;
      const float* v0 = vec0 + s * BLK;
      const float* v1 = vec1 + s * BLK;
      wait_bar(&split[s], (j / NS) & 1);        // vectors of this stage visible (long complete)
      wait_bar(&p1_ready[j & 1], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t ts = tw + Cfg::cSet + (j & 1) * Cfg::kSetCols;
      float alpha = 1.0f;
      if (MODE == M_FWD) {
        float mx = -INFINITY;
        if (!(p.dbg & 2))
#pragma unroll
        for (int c = 0; c < BLK / 32; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(ts + 32 * c, v);
          tmem_ld_wait();
#pragma unroll
          for (int t = 0; t < 32; ++t) mx = fmaxf(mx, __uint_as_float(v[t]) + v0[32 * c + t]);
        }
        const float m_new = fmaxf(m_run, mx);
        const float m_use = m_new == -INFINITY ? 0.0f : m_new;     // fully masked so far: keep exp2 finite
        alpha = ex2(m_run - m_use);
        float rs = 0.0f;
        if (!(p.dbg & 2))
#pragma unroll
        for (int c = 0; c < BLK / 32; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(ts + 32 * c, v);
          tmem_ld_wait();
          uint4 w4 = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
          for (int t = 0; t < 32; ++t) {
            float pv = ex2(__uint_as_float(v[t]) + v0[32 * c + t] - m_use);
            rs += pv;                                               // the row sum is taken before the dropout
            if (drop) {
              if ((t & 3) == 0) w4 = dropout_words(dseed, p.drop.site, qrow0 + i, (unsigned)(j * BLK + 32 * c + t) >> 2);
              pv = word_of(w4, t & 3) >= p.drop.thr ? pv * p.drop.inv_keep : 0.0f;
            }
            v[t] = __float_as_uint(pv);
          }
          tmem_st_32x32(ts + 32 * c, v);                          // P over S
#pragma unroll
          for (int t = 0; t < 32; ++t) v[t] = __float_as_uint(lo_comp<BLK / 8>(__uint_as_float(v[t]), 32 * c + t));
          tmem_st_32x32(ts + BLK + 32 * c, v);
        }
        l_run = l_run * alpha + rs;
        m_run = m_new;
      } else if (MODE == M_DQ) {
        if (!(p.dbg & 2))
#pragma unroll
        for (int c = 0; c < BLK / 32; ++c) {
          uint32_t sv[32], dp[32];
          tmem_ld_32x32(ts + 32 * c, sv);
          tmem_ld_32x32(ts + BLK + 32 * c, dp);
          tmem_ld_wait();
          uint4 w4 = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
          for (int t = 0; t < 32; ++t) {
            const float pv = ex2(__uint_as_float(sv[t]) + v0[32 * c + t] - lse2);
            float dpv = __uint_as_float(dp[t]);
            if (drop) {       // d(dropped P) -> dP: same mask, same 1/(1-p)
              if ((t & 3) == 0) w4 = dropout_words(dseed, p.drop.site, qrow0 + i, (unsigned)(j * BLK + 32 * c + t) >> 2);
              dpv = word_of(w4, t & 3) >= p.drop.thr ? dpv * p.drop.inv_keep : 0.0f;
            }
            const float ds = pv * (dpv - delta) * p.scale;
            dp[t] = __float_as_uint(ds);
            sv[t] = __float_as_uint(lo_comp<BLK / 8>(ds, 32 * c + t));
          }
          tmem_st_32x32(ts + BLK + 32 * c, dp);
          tmem_st_32x32(ts + 32 * c, sv);
        }
      } else {
        if (!(p.dbg & 2))
#pragma unroll
        for (int c = 0; c < BLK / 32; ++c) {
          uint32_t sv[32], dp[32];
          tmem_ld_32x32(ts + 32 * c, sv);
          tmem_ld_32x32(ts + 2 * BLK + 32 * c, dp);
          tmem_ld_wait();
#pragma unroll
          for (int t = 0; t < 32; ++t) {
            const float pv = valid ? ex2(__uint_as_float(sv[t]) - v0[32 * c + t]) : 0.0f;
            float dpv = __uint_as_float(dp[t]), pd = pv;
            if (drop) {       // rows are keys here: the mask word of (query row, this key) is word key % 4
              const uint4 w4 = dropout_words(dseed, p.drop.site, qrow0 + (unsigned)(j * BLK + 32 * c + t), (unsigned)i >> 2);
              const float kf = word_of(w4, i & 3) >= p.drop.thr ? p.drop.inv_keep : 0.0f;
              pd = pv * kf;
              dpv *= kf;
            }
            const float ds = pv * (dpv - v1[32 * c + t]) * p.scale;
            sv[t] = __float_as_uint(pd);
            dp[t] = __float_as_uint(ds);
          }
          tmem_st_32x32(ts + 32 * c, sv);                         // P^T over S^T
          tmem_st_32x32(ts + 2 * BLK + 32 * c, dp);               // dS^T over dP^T
#pragma unroll
          for (int t = 0; t < 32; ++t) {
            sv[t] = __float_as_uint(lo_comp<BLK / 8>(__uint_as_float(sv[t]), 32 * c + t));
            dp[t] = __float_as_uint(lo_comp<BLK / 8>(__uint_as_float(dp[t]), 32 * c + t));
          }
          tmem_st_32x32(ts + BLK + 32 * c, sv);
          tmem_st_32x32(ts + 3 * BLK + 32 * c, dp);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&a2_ready[j & 1]);
      if (j >= 1) consume(j - 1);
      alpha_pend = alpha;
    }
    consume(n_blk - 1);

    if (MODE == M_FWD) {
      if (valid) {
        const float inv = l_run > 0.0f ? 1.0f / l_run : 0.0f;
        store_row<HD>(p.o.p + b * p.o.sb + (long long)i * p.o.ld + h * HD, acc0, inv);
        p.lse[(long long)bh * p.Lq + i] = l_run > 0.0f ? m_run + log2f(l_run) : INFINITY;
      }
    } else if (MODE == M_DQ) {
      if (valid) store_row<HD>(p.dq.p + b * p.dq.sb + (long long)i * p.dq.ld + h * HD, acc0, 1.0f);
    } else if (i < p.Lk) {
      store_row<HD>(p.dv.p + b * p.dv.sb + (long long)i * p.dv.ld + h * HD, acc0, 1.0f);
      store_row<(MODE == M_DKV ? HD : 1)>(p.dk.p + b * p.dk.sb + (long long)i * p.dk.ld + h * HD, acc1, 1.0f);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc<Cfg::kCols>(tbase);
  }
}

// ------------------------------------------------------------------------------------------ host
using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                              const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                              CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeFn encode_fn() {
  static EncodeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn>(f);
  });
  return fn;
}

// [B, L, nh, hd] view (hd contiguous, row stride ld, batch stride sb, head stride hd) as a 4-D map
// {hd, L, nh, B}.  K-major use: boxes of 32 floats x box_rows rows (SWIZZLE_128B); MN-major use:
// 32 x 32 atoms (SWIZZLE_128B_ATOM_32B).  Rows beyond L are zero-filled.
static int make_map(CUtensorMap* tm, const float* ptr, long long ld, long long sb, int L, int nh, int hd, int B,
                    int box_rows, bool mn) {
  EncodeFn enc = encode_fn();
  if (!enc) return set_error(ITN_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[4] = {(cuuint64_t)hd, (cuuint64_t)L, (cuuint64_t)nh, (cuuint64_t)B};
  const cuuint64_t bstride = B > 1 ? (cuuint64_t)sb * 4 : (cuuint64_t)ld * 4 * (cuuint64_t)L;
  cuuint64_t gstr[3] = {(cuuint64_t)ld * 4, (cuuint64_t)hd * 4, bstride};
  cuuint32_t box[4] = {32u, mn ? 32u : (cuuint32_t)box_rows, 1u, 1u};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(ptr), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   mn ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(ITN_ERR_CUDA, "attention: cuTensorMapEncodeTiled failed (%d): dims %llu %llu %llu %llu "
                     "strides %llu %llu %llu", (int)r, gdim[0], gdim[1], gdim[2], gdim[3], gstr[0], gstr[1], gstr[2]);
  return ITN_OK;
}

static bool view_ok(const float* ptr, long long ld, long long sb, int B) {
  return ptr != nullptr && (reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && ld > 0 && (ld & 3) == 0 &&
         (B <= 1 || (sb > 0 && (sb & 3) == 0));
}

static int validate(const itn_attention_desc_t* d, bool bwd) {
  ITN_REQUIRE(d != nullptr, "attention: null descriptor");
  ITN_REQUIRE(d->B > 0 && d->nh > 0 && d->Lq > 0 && d->Lk > 0, "attention: B, nh, Lq, Lk must be positive");
  if (d->hd != 32 && d->hd != 64)
    return set_error(ITN_ERR_UNSUPPORTED, "attention: head dim %d (32 and 64 are built)", d->hd);
  ITN_REQUIRE((long long)d->B * d->nh * ((d->Lq > d->Lk ? d->Lq : d->Lk) / kRows + 1) < 0x7fffffffLL,
              "attention: grid too large");
  ITN_REQUIRE(d->lse != nullptr, "attention: lse is required");
  const bool ok = view_ok(d->q, d->q_ld, d->q_sb, d->B) && view_ok(d->k, d->k_ld, d->k_sb, d->B) &&
                  view_ok(d->v, d->v_ld, d->v_sb, d->B) && view_ok(d->o, d->o_ld, d->o_sb, d->B);
  if (!ok)
    return set_error(ITN_ERR_UNSUPPORTED, "attention: q/k/v/o must be 16-byte aligned with row/batch strides "
                     "multiples of 4 elements");
  if (bwd) {
    ITN_REQUIRE(d->delta != nullptr, "attention_bwd: delta scratch is required");
    const bool okb = view_ok(d->d_o, d->do_ld, d->do_sb, d->B) && view_ok(d->dq, d->dq_ld, d->dq_sb, d->B) &&
                     view_ok(d->dk, d->dk_ld, d->dk_sb, d->B) && view_ok(d->dv, d->dv_ld, d->dv_sb, d->B);
    if (!okb)
      return set_error(ITN_ERR_UNSUPPORTED, "attention_bwd: dO/dq/dk/dv must be 16-byte aligned with row/batch "
                       "strides multiples of 4 elements");
  }
  return ITN_OK;
}

static Params make_params(const itn_attention_desc_t* d, int rows) {
  Params p;
  auto rv = [](const float* ptr, long long ld, long long sb) {
    RowView v;
    v.p = const_cast<float*>(ptr);
    v.ld = ld;
    v.sb = sb;
    return v;
  };
  p.q = rv(d->q, d->q_ld, d->q_sb);
  p.k = rv(d->k, d->k_ld, d->k_sb);
  p.v = rv(d->v, d->v_ld, d->v_sb);
  p.o = rv(d->o, d->o_ld, d->o_sb);
  p.d_o = rv(d->d_o, d->do_ld, d->do_sb);
  p.dq = rv(d->dq, d->dq_ld, d->dq_sb);
  p.dk = rv(d->dk, d->dk_ld, d->dk_sb);
  p.dv = rv(d->dv, d->dv_ld, d->dv_sb);
  p.lse = d->lse;
  p.delta = d->delta;
  p.kmask = d->key_mask;
  p.nh = d->nh;
  p.Lq = d->Lq;
  p.Lk = d->Lk;
  p.tiles = (rows + kRows - 1) / kRows;
  p.scale = d->scale;
  p.scale_log2 = d->scale * 1.4426950408889634f;
  p.dbg = getenv("ITN_ATTN_DBG") ? atoi(getenv("ITN_ATTN_DBG")) : 0;
  p.drop.seed = d->drop_p > 0.f ? d->drop_seed : nullptr;
  p.drop.site = d->drop_site;
  p.drop.thr = (unsigned int)((double)d->drop_p * 4294967296.0);
  p.drop.inv_keep = d->drop_p > 0.f ? 1.0f / (1.0f - d->drop_p) : 1.0f;
  return p;
}

template <class K>
static int set_smem(K kern, int bytes, bool* done) {
  if (*done) return ITN_OK;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess)
    return set_error(ITN_ERR_CUDA, "attention: cudaFuncSetAttribute(%d B): %s", bytes, cudaGetErrorString(e));
  *done = true;
  return ITN_OK;
}

template <int HD, int BLK, int MINB>
static int launch_fwd(const itn_attention_desc_t* d, cudaStream_t s) {
  using Cfg = FwdCfg<HD, BLK>;
  Params p = make_params(d, d->Lq);
  CUtensorMap tmK, tmV;
  int rc = make_map(&tmK, d->k, d->k_ld, d->k_sb, d->Lk, d->nh, HD, d->B, BLK, false);
  if (rc) return rc;
  rc = make_map(&tmV, d->v, d->v_ld, d->v_sb, d->Lk, d->nh, HD, d->B, BLK, true);
  if (rc) return rc;
  auto kern = attn_fwd_kernel<HD, BLK, MINB>;
  static bool attr = false;
  rc = set_smem(kern, Cfg::kSmem, &attr);
  if (rc) return rc;
  launch(kern, d->B * d->nh * p.tiles, kThreads, Cfg::kSmem, s, tmK, tmV, p);
  return check_launch("attn_fwd_kernel");
}

template <int HD, int BLK, int MINB>
static int launch_dq(const itn_attention_desc_t* d, cudaStream_t s) {
  using Cfg = DqCfg<HD, BLK>;
  Params p = make_params(d, d->Lq);
  CUtensorMap tmKk, tmVk, tmKm;
  int rc = make_map(&tmKk, d->k, d->k_ld, d->k_sb, d->Lk, d->nh, HD, d->B, BLK, false);
  if (rc) return rc;
  rc = make_map(&tmVk, d->v, d->v_ld, d->v_sb, d->Lk, d->nh, HD, d->B, BLK, false);
  if (rc) return rc;
  rc = make_map(&tmKm, d->k, d->k_ld, d->k_sb, d->Lk, d->nh, HD, d->B, BLK, true);
  if (rc) return rc;
  auto kern = attn_bwd_dq_kernel<HD, BLK, MINB>;
  static bool attr = false;
  rc = set_smem(kern, Cfg::kSmem, &attr);
  if (rc) return rc;
  launch(kern, d->B * d->nh * p.tiles, kThreads, Cfg::kSmem, s, tmKk, tmVk, tmKm, p);
  return check_launch("attn_bwd_dq_kernel");
}

template <int HD, int BLK, int MINB, bool DROP>
static int launch_dkv_t(const itn_attention_desc_t* d, cudaStream_t s) {
  using Cfg = DkvCfg<HD, BLK>;
  Params p = make_params(d, d->Lk);
  CUtensorMap tmQk, tmDOk, tmQm, tmDOm;
  int rc = make_map(&tmQk, d->q, d->q_ld, d->q_sb, d->Lq, d->nh, HD, d->B, BLK, false);
  if (rc) return rc;
  rc = make_map(&tmDOk, d->d_o, d->do_ld, d->do_sb, d->Lq, d->nh, HD, d->B, BLK, false);
  if (rc) return rc;
  rc = make_map(&tmQm, d->q, d->q_ld, d->q_sb, d->Lq, d->nh, HD, d->B, BLK, true);
  if (rc) return rc;
  rc = make_map(&tmDOm, d->d_o, d->do_ld, d->do_sb, d->Lq, d->nh, HD, d->B, BLK, true);
  if (rc) return rc;
  auto kern = attn_bwd_dkv_kernel<HD, BLK, MINB, DROP>;
  static bool attr = false;
  rc = set_smem(kern, Cfg::kSmem, &attr);
  if (rc) return rc;
  launch(kern, d->B * d->nh * p.tiles, kThreads, Cfg::kSmem, s, tmQk, tmDOk, tmQm, tmDOm, p);
  return check_launch("attn_bwd_dkv_kernel");
}

template <int MODE, int HD, int BLK, int MINB, bool DROP>
static int launch_pipe_t(const itn_attention_desc_t* d, cudaStream_t s) {
  using Cfg = PipeCfg<MODE, HD, BLK>;
  Params p = make_params(d, MODE == M_DKV ? d->Lk : d->Lq);
  CUtensorMap tm[4];
  int rc;
  if (MODE == M_DKV) {
    rc = make_map(&tm[0], d->q, d->q_ld, d->q_sb, d->Lq, d->nh, HD, d->B, BLK, false);
    if (!rc) rc = make_map(&tm[1], d->d_o, d->do_ld, d->do_sb, d->Lq, d->nh, HD, d->B, BLK, false);
    if (!rc) rc = make_map(&tm[2], d->q, d->q_ld, d->q_sb, d->Lq, d->nh, HD, d->B, BLK, true);
    if (!rc) rc = make_map(&tm[3], d->d_o, d->do_ld, d->do_sb, d->Lq, d->nh, HD, d->B, BLK, true);
  } else if (MODE == M_DQ) {
    rc = make_map(&tm[0], d->k, d->k_ld, d->k_sb, d->Lk, d->nh, HD, d->B, BLK, false);
    if (!rc) rc = make_map(&tm[1], d->v, d->v_ld, d->v_sb, d->Lk, d->nh, HD, d->B, BLK, false);
    if (!rc) rc = make_map(&tm[2], d->k, d->k_ld, d->k_sb, d->Lk, d->nh, HD, d->B, BLK, true);
    tm[3] = tm[0];
  } else {
    rc = make_map(&tm[0], d->k, d->k_ld, d->k_sb, d->Lk, d->nh, HD, d->B, BLK, false);
    if (!rc) rc = make_map(&tm[1], d->v, d->v_ld, d->v_sb, d->Lk, d->nh, HD, d->B, BLK, true);
    tm[2] = tm[0];
    tm[3] = tm[0];
  }
  if (rc) return rc;
  auto kern = attn_pipe_kernel<MODE, HD, BLK, MINB, DROP>;
  static bool attr = false;
  rc = set_smem(kern, Cfg::kSmem, &attr);
  if (rc) return rc;
  launch(kern, d->B * d->nh * p.tiles, kPipeThreads, Cfg::kSmem, s, tm[0], tm[1], tm[2], tm[3], p);
  return check_launch("attn_pipe_kernel");
}

template <int HD, int BLK, int MINB>
static int launch_dkv(const itn_attention_desc_t* d, cudaStream_t s) {
  return d->drop_p > 0.f ? launch_dkv_t<HD, BLK, MINB, true>(d, s) : launch_dkv_t<HD, BLK, MINB, false>(d, s);
}

template <int MODE, int HD, int BLK, int MINB = 1>
static int launch_pipe(const itn_attention_desc_t* d, cudaStream_t s) {
  return d->drop_p > 0.f ? launch_pipe_t<MODE, HD, BLK, 1, true>(d, s) : launch_pipe_t<MODE, HD, BLK, MINB, false>(d, s);
}

static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

}  // namespace attn
}  // namespace itn

extern "C" int itn_attention_supported(const itn_attention_desc_t* d) {
  using namespace itn::attn;
  if (!d || d->B <= 0 || d->nh <= 0 || d->Lq <= 0 || d->Lk <= 0 || (d->hd != 32 && d->hd != 64)) return 0;
  return view_ok(d->q, d->q_ld, d->q_sb, d->B) && view_ok(d->k, d->k_ld, d->k_sb, d->B) &&
         view_ok(d->v, d->v_ld, d->v_sb, d->B) && view_ok(d->o, d->o_ld, d->o_sb, d->B);
}

// Which kernel runs a pass: measured on B200 (tools/attn_sweep.py, profiles/README.md).  "seq": the
// sequential-phase kernels (two CTAs per SM overlap each other); "pipe": the software-pipelined kernel.
// ITN_ATTN_FWD / ITN_ATTN_DQ / ITN_ATTN_DKV = seq | pipe and ITN_ATTN_*_BLK override for experiments.
static bool want_pipe(const char* name, bool dflt) {
  const char* v = getenv(name);
  if (!v) return dflt;
  return v[0] == 'p';
}

extern "C" int itn_attention_fwd(const itn_attention_desc_t* d, void* stream) {
  using namespace itn::attn;
  int rc = validate(d, false);
  if (rc) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int blk = env_int("ITN_ATTN_FWD_BLK", 0);
  const bool dropping = d->drop_p > 0.f;      // dropout lives in the pipelined kernels (and the hd-64 dK/dV kernel)
  if (dropping) ITN_REQUIRE(d->drop_seed != nullptr && d->drop_p < 1.f, "attention: dropout needs a seed and p < 1");
  if (dropping || want_pipe("ITN_ATTN_FWD", d->hd == 64)) {
    if (d->hd == 32) return blk == 32 ? launch_pipe<M_FWD, 32, 32, 2>(d, s) : launch_pipe<M_FWD, 32, 64>(d, s);
    return blk == 32 ? launch_pipe<M_FWD, 64, 32>(d, s) : launch_pipe<M_FWD, 64, 64>(d, s);
  }
  if (d->hd == 32) {
    if (blk == 32) return launch_fwd<32, 32, 2>(d, s);
    if (blk == 128) return launch_fwd<32, 128, 1>(d, s);
    return launch_fwd<32, 64, 2>(d, s);
  }
  if (blk == 64) return launch_fwd<64, 64, 1>(d, s);
  return launch_fwd<64, 32, 2>(d, s);
}

extern "C" int itn_attention_bwd(const itn_attention_desc_t* d, void* stream) {
  using namespace itn::attn;
  int rc = validate(d, true);
  if (rc) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int bq = env_int("ITN_ATTN_DQ_BLK", 0), bkv = env_int("ITN_ATTN_DKV_BLK", 0);
  const int only = env_int("ITN_ATTN_BWD_ONLY", 0);   // timing experiments: 1 = dQ kernel only, 2 = dK/dV kernel only
  // dQ first: it also writes delta = rowsum(dO * O), which the dK/dV kernel reads
  if (only != 2) {
    const bool dropping = d->drop_p > 0.f;
    if (dropping) ITN_REQUIRE(d->drop_seed != nullptr && d->drop_p < 1.f, "attention: dropout needs a seed and p < 1");
    if (dropping || want_pipe("ITN_ATTN_DQ", false)) {
      if (d->hd == 32) rc = bq == 32 ? launch_pipe<M_DQ, 32, 32>(d, s) : launch_pipe<M_DQ, 32, 64>(d, s);
      else rc = launch_pipe<M_DQ, 64, 32>(d, s);
    } else if (d->hd == 32) {
      if (bq == 64) rc = launch_dq<32, 64, 1>(d, s);
      else if (bq == 128) rc = launch_dq<32, 128, 1>(d, s);
      else rc = launch_dq<32, 32, 2>(d, s);
    } else {
      if (bq == 32) rc = launch_dq<64, 32, 1>(d, s);
      else rc = launch_dq<64, 64, 1>(d, s);
    }
    if (rc) return rc;
  }
  if (only == 1) return ITN_OK;
  if (d->hd == 32) {
    if (d->drop_p > 0.f || want_pipe("ITN_ATTN_DKV", true)) return launch_pipe<M_DKV, 32, 32>(d, s);
    if (bkv == 32) return launch_dkv<32, 32, 1>(d, s);
    return launch_dkv<32, 64, 1>(d, s);
  }
  return launch_dkv<64, 32, 1>(d, s);
}

// ---- tensor-pipe micro-benchmark (tools/mma_bench.py): cycles per tcgen05.mma kind::tf32 128 x N x 8 for the
// operand sources / layouts the attention kernels use.  One warp per CTA; lane 0 issues `iters` MMAs into one
// accumulator, commits and waits; out[blockIdx.x] = elapsed cycles.  Operand contents are whatever shared /
// tensor memory holds (timing only).
namespace itn {
namespace attn {
template <int N>
__global__ void __launch_bounds__(32, 1) mma_bench_kernel(int a_tmem, int b_mn, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 96 * 1024);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 1);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tmem_ptr;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 32) reinterpret_cast<float*>(smem)[i] = 1.0f + i * 1e-6f;
  fence_proxy_async();
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t sa = smem_u32(smem), sb = smem_u32(smem + 32 * 1024);
    const uint32_t idesc = umma_idesc_tf32(kRows, N, 0, b_mn ? 1 : 0);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const int ks = i & 3;
      const uint64_t bd = b_mn ? umma_smem_desc(sb + ks * 1024, 4096, 512, kLayoutSW128Base32)
                               : umma_smem_desc(sb + ks * 32, 16, 1024, kLayoutSW128);
      if (a_tmem) umma_tf32_ts(tbase, tbase + 256 + 8 * ks, bd, idesc, i > 0 ? 1u : 0u);
      else umma_tf32(tbase, umma_smem_desc(sa + ks * 32, 16, 1024, kLayoutSW128), bd, idesc, i > 0 ? 1u : 0u);
    }
    const long long t1 = clock64();
    umma_commit(bar);
    wait_bar(bar, 0);
    const long long t2 = clock64();
    out[2 * blockIdx.x] = t2 - t0;
    out[2 * blockIdx.x + 1] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  tmem_dealloc<512>(tbase);
}
}  // namespace attn
}  // namespace itn

extern "C" int itn_debug_mma_bench(int n, int a_tmem, int b_mn, int iters, int grid, long long* out, void* stream) {
  using namespace itn::attn;
  const int smem = 96 * 1024 + 64 + 1024;
  void (*kern)(int, int, int, long long*) = nullptr;
  switch (n) {
    case 32: kern = mma_bench_kernel<32>; break;
    case 64: kern = mma_bench_kernel<64>; break;
    case 128: kern = mma_bench_kernel<128>; break;
    case 256: kern = mma_bench_kernel<256>; break;
    default: return itn::set_error(ITN_ERR_ARG, "mma_bench: N must be 32, 64, 128 or 256");
  }
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return itn::set_error(ITN_ERR_CUDA, "mma_bench: %s", cudaGetErrorString(e));
  itn::launch(kern, grid, 32, smem, static_cast<cudaStream_t>(stream), a_tmem, b_mn, iters, out);
  return itn::check_launch("mma_bench_kernel");
}
