// Batched TF32 GEMM on tcgen05 tensor cores (sm_100a), operands staged by TMA,
// accumulators in TMEM, fused epilogue.  See include/interactron_b200.h
// (itn_gemm_tf32) for the contract and the reference call sites it replaces.
//
// Persistent, warp-specialised kernel: one CTA per SM loops over the 128 x BN output tiles
// (n fastest, so concurrently running CTAs share A tiles in L2).  Roles:
//   warp 0      TMA producer: fills a STAGES-deep ring of {A tile, B tile} with 128-byte
//               swizzled boxes whose inner extent is 32 floats; runs ahead across tiles.
//   warp 1      TMEM owner + MMA issuer: one lane issues tcgen05.mma.kind::tf32 128xBNx8 into
//               one of TWO TMEM accumulators, so tile i+1 is computed while tile i drains.
//   warps 2-5   (tf32x3 mode only) residual splitters, see "Precision modes".
//   last 4      epilogue: TMEM -> registers -> per-warp smem transpose -> coalesced global
//               stores, with bias / activation / mask / residual / accumulate fused and the
//               auxiliary global loads of a row block issued together (latency hidden).
// Both operand majors are supported through the UMMA descriptors, so forward (x W^T),
// data-grad (dy W) and weight-grad (dy^T x) all run without materialising transposes.
// Out-of-bounds rows/cols/K are zero-filled by TMA and masked in the epilogue,
// so ragged sizes (N=1236, K=1496, M=250, ...) need no padding in HBM.
//
// Precision modes.  kind::tf32 truncates its fp32 operands to 10 mantissa bits; one pass
// ("tf32") therefore carries ~1e-3 error per GEMM, which the inner-loop gradient amplifies to
// 1-10 % (measured).  The default "tf32x3" mode is error-compensated: while the
// raw tiles sit in shared memory, four splitter warps write the residual tiles
// x_lo = x - trunc_tf32(x)  next to them, and the issuer accumulates
//     A*B  +  A_lo*B  +  A*B_lo        (the tensor core sees A, B as their truncated hi parts)
// in TMEM, which restores ~fp32 accuracy (dropped terms are O(2^-20)) at 3 MMAs per k-step
// and no extra HBM/L2 traffic.
#include <cuda.h>
#include <cuda_runtime.h>

#include <mutex>

#include "itn_common.cuh"
#include "itn_ptx.cuh"

namespace itn {

// Optional in-kernel timeline (debug builds only: `make trace`): CTA 0 stamps clock64() at the
// hand-off points of every role so pipeline bubbles can be read off directly.
#ifdef ITN_TRACE
__device__ long long* g_trace = nullptr;
#define ITN_TRACE_AT(slot, idx)                                                        \
  do {                                                                                 \
    if (blockIdx.x == 0 && g_trace != nullptr && (idx) < 1024) g_trace[(slot) * 1024 + (idx)] = clock64(); \
  } while (0)
#else
#define ITN_TRACE_AT(slot, idx) do { } while (0)
#endif

// x - trunc_tf32(x): the part of an fp32 operand the tensor core drops (kind::tf32 truncates).  Exact in fp32
// (13 significant bits); the tensor core truncates it to 11.  Rounding it to nearest TF32 here (as the attention
// kernels do, tf32_lo) was measured: small-K products 6.3e-7 -> 4.7e-7 vs fp64, but the splitter warps are the
// critical path of the tf32x3 main loop and the two extra integer ops per element cost 16480x2048x512
// 221 -> 186 TFLOP/s and the step 8 % (cvt.rna: 163 TFLOP/s).  Not worth it: left exact.
__device__ __forceinline__ float4 tf32_residual(const float4 v) {
  return make_float4(tf32_lo_exact(v.x), tf32_lo_exact(v.y), tf32_lo_exact(v.z), tf32_lo_exact(v.w));
}

#ifndef ITN_GEMM_PAIR_DEFAULT
#define ITN_GEMM_PAIR_DEFAULT 1
#endif
constexpr int kBM = 128;
#ifndef ITN_BK
#define ITN_BK 32
#endif
constexpr int kBK = ITN_BK;                // floats per k-block: 32 (128 B swizzle rows) or 16 (64 B rows, twice the stages)
static_assert(kBK == 32 || kBK == 16, "k-block of 16 or 32 floats");
constexpr int kAtomBytes = 32 * kBK * 4;   // one 32(mn) x 32(k) MN-major box = 4096 B

struct GemmKParams {
  int M, N, K;
  int nb1;
  int tiles_m, tiles_n, num_tiles;
  int a_m0, a_m1, b_m0, b_m1;  // 0 when the operand is broadcast over that batch dim
  float* C;              long long ldc,   c_sb0,    c_sb1;
  const float* bias;     long long        bias_sb0, bias_sb1;
  const float* residual; long long ldr,   r_sb0,    r_sb1;
  const float* aux;      long long ldaux, aux_sb0,  aux_sb1;
  float* C2;             long long ldc2,  c2_sb0,   c2_sb1;
  float alpha;
  int act, epi, accumulate, round_out, act_pos;
  int variant;   // epilogue_variant(...)
  int vec;       // 1: 128-bit epilogue path is legal (alignment / N % 4 checked on the host)
  int split_acc; // tf32x3, tiles <= 128 wide: residual products A_lo*B + A*B_lo accumulate in their own TMEM columns and
                 // are added in the epilogue - the main accumulator sees K/8 round-toward-zero accumulates, not 3K/8
  float rz_c0;   // tf32x3: chain-length independent part of the shrink (truncation of the residual operands, dropped lo*lo)
  float rz_eps;  // tf32x3: mean relative loss of one round-toward-zero TMEM accumulate (see splitters)
  // implicit-GEMM convolution (conv_kw > 0): A tiles come through an im2col tensor map
  int conv_kh, conv_kw, conv_c, conv_stride, conv_pad, conv_dil, conv_wo, conv_howo;
  int b_presplit;  // tf32x3, K-major B: the residual tile of B is loaded through tmBlo (itn_gemm_desc_t::B_lo)
  int dbg;       // ITN_TRACE builds only: 1 = skip global stores, 2 = skip TMEM read, 4 = skip smem transpose;
                 // any build (ITN_GEMM_DBG, timing experiments): 16 = skip the residual split, 32 = skip the correction MMAs
};

// Per-batch-entry epilogue pointers.
struct EpiPtrs {
  float* C;
  const float* bias;
  const float* residual;
  const float* aux;
  float* C2;
};

__device__ __forceinline__ EpiPtrs make_epi_ptrs(const GemmKParams& p, int b0, int b1) {
  EpiPtrs e;
  e.C = p.C + b0 * p.c_sb0 + b1 * p.c_sb1;
  e.bias = p.bias ? p.bias + b0 * p.bias_sb0 + b1 * p.bias_sb1 : nullptr;
  e.residual = p.residual ? p.residual + b0 * p.r_sb0 + b1 * p.r_sb1 : nullptr;
  e.aux = p.aux ? p.aux + b0 * p.aux_sb0 + b1 * p.aux_sb1 : nullptr;
  e.C2 = p.C2 ? p.C2 + b0 * p.c2_sb0 + b1 * p.c2_sb1 : nullptr;
  return e;
}

__device__ __forceinline__ void epilogue_store(const GemmKParams& p, const EpiPtrs& e, int row,
                                               int col, float acc, float bias_v) {
  float v = p.alpha * acc + bias_v;
  if (e.C2) e.C2[(long long)row * p.ldc2 + col] = v;
  if (p.act_pos == 0) {
    if (p.act == ITN_ACT_RELU) v = fmaxf(v, 0.0f);
    else if (p.act == ITN_ACT_GELU) v = gelu_erf(v);
  }
  if (p.epi == ITN_EPI_RELU_MASK) {
    v = e.aux[(long long)row * p.ldaux + col] > 0.0f ? v : 0.0f;
  } else if (p.epi == ITN_EPI_GELU_GRAD) {
    v *= gelu_erf_grad(e.aux[(long long)row * p.ldaux + col]);
  }
  if (e.residual) v += e.residual[(long long)row * p.ldr + col];
  float* c = e.C + (long long)row * p.ldc + col;
  if (p.accumulate) v += *c;
  if (p.act_pos == 1) {
    if (p.act == ITN_ACT_RELU) v = fmaxf(v, 0.0f);
    else if (p.act == ITN_ACT_GELU) v = gelu_erf(v);
  }
  *c = p.round_out ? rn_tf32(v) : v;
}

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == ITN_ACT_RELU) return fmaxf(v, 0.0f);
  if (act == ITN_ACT_GELU) return gelu_erf(v);
  return v;
}

// Which operand of the general epilogue is streamed through the register prefetch.
enum { XS_NONE = 0, XS_AUX = 1, XS_RESIDUAL = 2, XS_CIN = 3 };

// Epilogue variants.  Every fused combination the path uses gets its own branch-free,
// compile-time specialised row loop (a single generic loop with run-time flags measured 6-7x
// slower: the 32-row unrolled body with a dozen warp-uniform branches per element thrashes the
// instruction cache).  EV_GENERIC keeps the full contract for anything else.
enum {
  EV_SIMPLE = 0,        // bias + {none, relu, gelu}
  EV_RESIDUAL,          // + residual
  EV_RESIDUAL_RELU,     // relu(bias + acc + residual)            (ResNet bottleneck output)
  EV_ACCUMULATE,        // C += acc                               (gradient accumulation)
  EV_RELU_MASK,         // acc * (aux > 0)                        (d ReLU)
  EV_GELU_GRAD,         // acc * gelu'(aux)                       (d GELU)
  EV_PRE,               // C = C2 = bias + acc                    (rounded + full copies)
  EV_PRE_GELU,          // C2 = bias + acc, C = gelu(C2)
  EV_GENERIC
};

__host__ __device__ inline int epilogue_variant(bool aux, bool residual, bool c2, int act, int epi,
                                                int accumulate, int act_pos) {
  const int extras = (aux ? 1 : 0) + (residual ? 1 : 0) + (accumulate ? 1 : 0);
  if (extras == 0 && !c2) return EV_SIMPLE;
  if (extras == 0 && c2 && epi == ITN_EPI_NONE) {
    if (act == ITN_ACT_NONE) return EV_PRE;
    if (act == ITN_ACT_GELU && act_pos == 0) return EV_PRE_GELU;
    return EV_GENERIC;
  }
  if (c2 || extras != 1) return EV_GENERIC;
  if (residual && epi == ITN_EPI_NONE) {
    if (act == ITN_ACT_NONE) return EV_RESIDUAL;
    if (act == ITN_ACT_RELU && act_pos == 1) return EV_RESIDUAL_RELU;
    return EV_GENERIC;
  }
  if (accumulate && epi == ITN_EPI_NONE && act == ITN_ACT_NONE) return EV_ACCUMULATE;
  if (aux && act == ITN_ACT_NONE) {
    if (epi == ITN_EPI_RELU_MASK) return EV_RELU_MASK;
    if (epi == ITN_EPI_GELU_GRAD) return EV_GELU_GRAD;
  }
  return EV_GENERIC;
}

// One transposed 32x32 accumulator chunk (st[r*33 + lane] = row r, this lane's column) -> global
// memory.  `xs` holds the streamed operand of this chunk (prefetched by the caller).
// Scalar (32-bit) variants for outputs whose rows are not 16-byte aligned or N % 4 != 0 (e.g. the
// attention score matrices): compact loops, the streamed operand is read in place.
template <int EV>
__device__ __forceinline__ void epilogue_rows(const GemmKParams& p, const EpiPtrs& e, const float* st,
                                              int lane, int row_base, int rmax, int col,
                                              const float* xsp, long long ldx) {
  const float bias_v = e.bias ? e.bias[col] : 0.0f;
  const float alpha = p.alpha;
  const bool rnd = p.round_out != 0;
  const long long ldc = p.ldc, ldc2 = p.ldc2;
  float* cp = e.C + (long long)row_base * ldc + col;
  float* c2p = (EV == EV_PRE || EV == EV_PRE_GELU) ? e.C2 + (long long)row_base * ldc2 + col : nullptr;
  const float* xp = xsp ? xsp + (long long)row_base * ldx + col : nullptr;
  const float* sp = st + lane;
#pragma unroll 4
  for (int r = 0; r < rmax; ++r) {
    float v = fmaf(alpha, sp[r * 33], bias_v);
    if (EV == EV_PRE || EV == EV_PRE_GELU) c2p[r * ldc2] = v;
    if (EV == EV_PRE_GELU) v = gelu_erf(v);
    if (EV == EV_RESIDUAL || EV == EV_RESIDUAL_RELU || EV == EV_ACCUMULATE) v += xp[r * ldx];
    if (EV == EV_RESIDUAL_RELU) v = fmaxf(v, 0.0f);
    if (EV == EV_RELU_MASK) v = xp[r * ldx] > 0.0f ? v : 0.0f;
    if (EV == EV_GELU_GRAD) v *= gelu_erf_grad(xp[r * ldx]);
    cp[r * ldc] = rnd ? rn_tf32(v) : v;
  }
}

__device__ __forceinline__ void epilogue_simple(const GemmKParams& p, const EpiPtrs& e, const float* st,
                                                int lane, int row_base, int rmax, int col) {
  const float bias_v = e.bias ? e.bias[col] : 0.0f;
  float* cp = e.C + (long long)row_base * p.ldc + col;
  const float alpha = p.alpha;
  const int act = p.act;
  const bool rnd = p.round_out != 0;
  // act / rounding are hoisted out of the row loops (warp-uniform)
  if (act == ITN_ACT_GELU) {
#pragma unroll 4
    for (int r = 0; r < rmax; ++r) {
      const float v = gelu_erf(fmaf(alpha, st[r * 33 + lane], bias_v));
      cp[(long long)r * p.ldc] = rnd ? rn_tf32(v) : v;
    }
  } else if (rmax == 32 && !rnd) {
    const float lo = act == ITN_ACT_RELU ? 0.0f : -INFINITY;
#pragma unroll
    for (int r = 0; r < 32; ++r)
      cp[(long long)r * p.ldc] = fmaxf(fmaf(alpha, st[r * 33 + lane], bias_v), lo);
  } else {
    const float lo = act == ITN_ACT_RELU ? 0.0f : -INFINITY;
#pragma unroll 4
    for (int r = 0; r < rmax; ++r) {
      const float v = fmaxf(fmaf(alpha, st[r * 33 + lane], bias_v), lo);
      cp[(long long)r * p.ldc] = rnd ? rn_tf32(v) : v;
    }
  }
}

// ---- 128-bit variants.  Staging pitch 36 floats: thread t stores its accumulator row t as 8 float4
// (conflict-free per quarter-warp), then lane (q = lane/8, c = lane%8) owns columns 4c..4c+3 of rows
// 4i+q: every warp-wide global access covers 4 rows x 128 contiguous bytes.  Needs 16-byte aligned
// rows (ld % 4 == 0) and N % 4 == 0 for C and the streamed operand (GemmKParams::vec).
__device__ __forceinline__ float4 f4_fma(float a, float4 x, float4 b) {
  return make_float4(fmaf(a, x.x, b.x), fmaf(a, x.y, b.y), fmaf(a, x.z, b.z), fmaf(a, x.w, b.w));
}
__device__ __forceinline__ float4 f4_add(float4 a, float4 b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
__device__ __forceinline__ float4 f4_max(float4 a, float lo) {
  return make_float4(fmaxf(a.x, lo), fmaxf(a.y, lo), fmaxf(a.z, lo), fmaxf(a.w, lo));
}
__device__ __forceinline__ float4 f4_rnd(float4 a) {
  return make_float4(rn_tf32(a.x), rn_tf32(a.y), rn_tf32(a.z), rn_tf32(a.w));
}

template <int EV, bool FULL, bool RND>
__device__ __forceinline__ void epilogue_rows_v4(const GemmKParams& p, const EpiPtrs& e, const float4* st4,
                                                 int lane, int row_base, int rmax, int col,
                                                 const float4 (&xs)[8]) {
  const int q = lane >> 3, c = lane & 7;
  const float4 bias_v = e.bias ? *reinterpret_cast<const float4*>(e.bias + col) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float alpha = p.alpha;
  const int act = p.act;
  const float lo = (EV == EV_SIMPLE && act == ITN_ACT_RELU) ? 0.0f : -INFINITY;
  const bool gelu = EV == EV_PRE_GELU || (EV == EV_SIMPLE && act == ITN_ACT_GELU);
  const long long step = 4 * p.ldc, step2 = 4 * p.ldc2;
  float* cp = e.C + (long long)(row_base + q) * p.ldc + col;
#ifdef ITN_TRACE
  // dbg 8: same instruction stream, but every chunk overwrites one L2-resident 16 KB window per CTA
  if (p.dbg & 8) cp = e.C + (long long)(blockIdx.x * 128 + (row_base & 127) + q) * p.ldc + (col & 31);
#endif
  float* c2p = (EV == EV_PRE || EV == EV_PRE_GELU) ? e.C2 + (long long)(row_base + q) * p.ldc2 + col : nullptr;
  const float4* sp = st4 + q * 9 + c;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (FULL || 4 * i + q < rmax) {
      float4 v = f4_fma(alpha, sp[i * 36], bias_v);
      if (EV == EV_PRE || EV == EV_PRE_GELU) *reinterpret_cast<float4*>(c2p + i * step2) = v;
      if (gelu) v = make_float4(gelu_erf(v.x), gelu_erf(v.y), gelu_erf(v.z), gelu_erf(v.w));
      if (EV == EV_SIMPLE) v = f4_max(v, lo);
      if (EV == EV_RESIDUAL || EV == EV_RESIDUAL_RELU || EV == EV_ACCUMULATE) v = f4_add(v, xs[i]);
      if (EV == EV_RESIDUAL_RELU) v = f4_max(v, 0.0f);
      if (EV == EV_RELU_MASK)
        v = make_float4(xs[i].x > 0.f ? v.x : 0.f, xs[i].y > 0.f ? v.y : 0.f, xs[i].z > 0.f ? v.z : 0.f,
                        xs[i].w > 0.f ? v.w : 0.f);
      if (EV == EV_GELU_GRAD)
        v = make_float4(v.x * gelu_erf_grad(xs[i].x), v.y * gelu_erf_grad(xs[i].y),
                        v.z * gelu_erf_grad(xs[i].z), v.w * gelu_erf_grad(xs[i].w));
      *reinterpret_cast<float4*>(cp + i * step) = RND ? f4_rnd(v) : v;
    }
  }
}

// warp-uniform dispatch on (full row block, rounding) so the common case has no per-row predicates
template <int EV>
__device__ __forceinline__ void epilogue_v4(const GemmKParams& p, const EpiPtrs& e, const float4* st4, int lane,
                                            int row_base, int rmax, int col, const float4 (&xs)[8]) {
  if (p.round_out) epilogue_rows_v4<EV, false, true>(p, e, st4, lane, row_base, rmax, col, xs);
  else if (rmax == 32) epilogue_rows_v4<EV, true, false>(p, e, st4, lane, row_base, rmax, col, xs);
  else epilogue_rows_v4<EV, false, false>(p, e, st4, lane, row_base, rmax, col, xs);
}

// Full contract with run-time flags; compact (not unrolled) on purpose.  Rarely taken.
__device__ __forceinline__ void epilogue_generic(const GemmKParams& p, const EpiPtrs& e, const float* st,
                                              int lane, int row_base, int rmax, int col) {
  const float bias_v = e.bias ? e.bias[col] : 0.0f;
#pragma unroll 1
  for (int r = 0; r < rmax; ++r)
    epilogue_store(p, e, row_base + r, col, st[r * 33 + lane], bias_v);
}

// PAIR: the CTA-pair (cta_group::2) variant - 256 x BN tiles computed by two CTAs of a cluster, each holding its
// 128 rows of A and HALF of the B tile (BN/2 rows): per CTA the shared-memory traffic of the B operand halves
// (TMA writes and MMA reads), which is what bounds the tf32x3 main loop, and a third stage fits.
template <int BN, bool X3, bool PAIR = false>
struct TileCfg {
  static constexpr int kABytes = kBM * kBK * 4;
  static constexpr int kBRows = PAIR ? BN / 2 : BN;            // B rows held by this CTA
  static constexpr int kBBytes = kBRows * kBK * 4;
  static constexpr int kRawBytes = kABytes + kBBytes;          // what TMA delivers per stage
  static constexpr int kStageBytes = X3 ? 2 * kRawBytes : kRawBytes;   // + residual (lo) tiles
  static constexpr int kStagingBytes = 4 * 32 * 36 * 4;        // per-warp transpose buffers (pitch 36)
  static constexpr int kBudget = 227 * 1024 - kStagingBytes - 1024 - 512;
  static constexpr int kMaxStages = kBudget / kStageBytes;
  static constexpr int kStageCap = kBK == 32 ? 8 : 16;
  static constexpr int kStages = kMaxStages > kStageCap ? kStageCap : kMaxStages;
  static constexpr int kAccCols = BN < 32 ? 32 : BN;
  // two accumulators (double buffer); tiles up to 128 wide also hold a second pair for the residual products
  // (GemmKParams::split_acc), 256-wide tiles fill tensor memory with the main pair alone
  static constexpr bool kCanSplit = BN <= 128;
  static constexpr int kTmemCols = kCanSplit ? 4 * kAccCols : 2 * kAccCols;
  static constexpr int kThreads = X3 ? 320 : 192;
  static constexpr int kEpiWarp0 = X3 ? 6 : 2;
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + (3 * kStages + 4) * 8 + 16 + 1024;
  static_assert(kStages >= 2, "need at least a double-buffered operand ring");
  static_assert(kSmemBytes <= 227 * 1024, "exceeds shared memory per CTA");
  static_assert(kSmemBytes > 114 * 1024, "one CTA per SM is assumed (TMEM is allocated in full)");
  static_assert(!PAIR || (X3 && BN == 256 && kBK == 32), "the CTA-pair variant is built for tf32x3 128x256x32 tiles");
};

// a protocol error in the pair kernel must not hang the GPU: trap after ~1 s of failed waits
__device__ __forceinline__ void mbar_wait_wd(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 25)) __trap();
  }
}

template <int BN, bool A_MN, bool B_MN, bool X3, bool PAIR = false>
__global__ void __launch_bounds__(TileCfg<BN, X3, PAIR>::kThreads, 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmBlo, const __grid_constant__ GemmKParams p) {
  // p is __grid_constant__: it is read straight from the constant bank even where its address is
  // taken (a by-value copy lands on the local-memory stack and turns every p.ld* into an LDL).
  using Cfg = TileCfg<BN, X3, PAIR>;
  static_assert(!PAIR || (!A_MN && !B_MN), "the CTA-pair variant takes K-major operands");
  // PAIR: CTA `rank` of the pair owns m-tile 2*pair_m + rank of the pair's tile and B rows [rank*BN/2, +BN/2);
  // the leader (rank 0) issues the MMAs for both, every barrier the issuer waits on collects both CTAs
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const int cta_id = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;       // tile-walk index
  const int cta_n = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  constexpr int STAGES = Cfg::kStages;

  extern __shared__ uint8_t smem_raw[];
  // 128-byte swizzle atoms need 1024-byte aligned tiles.  The alignment is applied as an OFFSET
  // on the __shared__ array so the compiler keeps the shared address space (rounding through
  // uintptr_t turns every staging / splitter access into a generic LD/ST: measured 4x slower).
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* staging = reinterpret_cast<float*>(smem + STAGES * Cfg::kStageBytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::kStageBytes + Cfg::kStagingBytes);
  uint64_t* split_bar = full_bar + STAGES;
  uint64_t* empty_bar = split_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;      // [2] accumulator ready for the epilogue
  uint64_t* tempty_bar = tfull_bar + 2;          // [2] accumulator drained
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kb = (p.K + kBK - 1) / kBK;
  if (threadIdx.x == 0) ITN_TRACE_AT(8, 8);    // CTA entry

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.b_presplit) tma_prefetch_desc(&tmBlo);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&split_bar[s], PAIR ? 8 : 4);   // one arrival per splitter warp (of both CTAs: the leader issues)
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], PAIR ? 8 : 4);  // one arrival per epilogue warp (of both CTAs)
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    if (PAIR) tmem_alloc_pair<Cfg::kTmemCols>(tmem_ptr_smem);
    else tmem_alloc<Cfg::kTmemCols>(tmem_ptr_smem);
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();      // the peer's barriers are initialised before anyone signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // Everything above touches only this CTA's shared memory / TMEM and overlaps the tail of the
  // previous kernel in the stream; global memory is first read (TMA) or written below.
  if (threadIdx.x == 0) ITN_TRACE_AT(8, 9);    // prologue done
  pdl_wait();
  pdl_trigger();

  if (warp == 0) {
    // ------------------------------------------------------- TMA producer (converged warp, elect.sync)
    {
      int s = 0;
      uint32_t ph = 0;
      int gk = 0;   // running k-block counter (trace index)
      for (int tile = cta_id; tile < p.num_tiles; tile += cta_n) {
        const int nb = tile % p.tiles_n;
        const int t2 = tile / p.tiles_n;
        const int mb = PAIR ? 2 * (t2 % p.tiles_m) + (int)rank : t2 % p.tiles_m;     // PAIR: tiles_m counts pairs
        const int z = t2 / p.tiles_m;
        const int b0 = z / p.nb1, b1 = z % p.nb1;
        const int m0 = mb * kBM, n0 = nb * BN + (PAIR ? (int)rank * (BN / 2) : 0);
        const int a_c0 = b0 * p.a_m0, a_c1 = b1 * p.a_m1;
        const int b_c0 = b0 * p.b_m0, b_c1 = b1 * p.b_m1;
        for (int kb = 0; kb < num_kb; ++kb, ++gk) {
          if (PAIR) mbar_wait_wd(&empty_bar[s], ph ^ 1); else mbar_wait(&empty_bar[s], ph ^ 1);
          if (elect_one()) {
            ITN_TRACE_AT(0, gk);
            // pre-split B (static weights): its residual tile arrives by TMA, the splitters do A only
            const bool presplit = X3 && !B_MN && p.b_presplit;
            mbar_expect_tx(&full_bar[s], Cfg::kRawBytes + (presplit ? Cfg::kBBytes : 0));
            uint8_t* sa = smem + s * Cfg::kStageBytes;
            uint8_t* sb = sa + Cfg::kABytes;
            const int k0 = kb * kBK;
            if (!A_MN && p.conv_kw > 0 && p.conv_c == 4) {
              // 4-channel input (the RGB stem, padded): one k-block = 8 filter taps x 4 channels.  Each tap is its own
              // im2col load of [128 pixels x 16 bytes] into a 2 KB slab: unswizzled K-major core matrices (8 rows x 16 B),
              // slabs = consecutive 16-byte K chunks (descriptor: LBO = slab, SBO = 128 B).  Taps beyond kh*kw re-read
              // the last tap; their weights are the zero-filled K tail of B.
              const int n = m0 / p.conv_howo, rem = m0 - n * p.conv_howo;
              const int py = rem / p.conv_wo, px = rem - py * p.conv_wo;
              const int taps = p.conv_kh * p.conv_kw;
#pragma unroll
              for (int j = 0; j < kBK / 4; ++j) {
                int tap = k0 / 4 + j;
                tap = tap < taps ? tap : taps - 1;
                const int ky = tap / p.conv_kw, kx = tap - ky * p.conv_kw;
                tma_load_im2col_4d(sa + j * (kBM * 16), &tmA, &full_bar[s], 0, px * p.conv_stride - p.conv_pad,
                                   py * p.conv_stride - p.conv_pad, n, (unsigned short)(kx * p.conv_dil),
                                   (unsigned short)(ky * p.conv_dil));
              }
            } else if (!A_MN && p.conv_kw > 0) {
              // k-block -> filter tap (ky, kx) and first channel; tile row m0 -> output pixel (n, py, px)
              const int tap = k0 / p.conv_c, c0 = k0 - tap * p.conv_c;
              const int ky = tap / p.conv_kw, kx = tap - ky * p.conv_kw;
              const int n = m0 / p.conv_howo, rem = m0 - n * p.conv_howo;
              const int py = rem / p.conv_wo, px = rem - py * p.conv_wo;
              tma_load_im2col_4d(sa, &tmA, &full_bar[s], c0, px * p.conv_stride - p.conv_pad,
                                 py * p.conv_stride - p.conv_pad, n, (unsigned short)(kx * p.conv_dil),
                                 (unsigned short)(ky * p.conv_dil));
            } else if (!A_MN) {
              tma_load_4d(sa, &tmA, &full_bar[s], k0, m0, a_c1, a_c0);
            } else {
#pragma unroll
              for (int j = 0; j < kBM / 32; ++j)
                tma_load_4d(sa + j * kAtomBytes, &tmA, &full_bar[s], m0 + 32 * j, k0, a_c1, a_c0);
            }
            if (!B_MN) {
              tma_load_4d(sb, &tmB, &full_bar[s], k0, n0, b_c1, b_c0);
              if (presplit) tma_load_4d(sb + Cfg::kRawBytes, &tmBlo, &full_bar[s], k0, n0, b_c1, b_c0);
            } else {
#pragma unroll
              for (int j = 0; j < BN / 32; ++j)
                tma_load_4d(sb + j * kAtomBytes, &tmB, &full_bar[s], n0 + 32 * j, k0, b_c1, b_c0);
            }
          }
          __syncwarp();
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // --------------------------------------------------------- MMA issuer (converged warp, elect.sync)
    if (!PAIR || rank == 0) {
      constexpr uint32_t idesc = umma_idesc_tf32(PAIR ? 2 * kBM : kBM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      int s = 0;
      uint32_t ph = 0;
      int acc = 0;
      uint32_t acc_ph = 0;
      int gk = 0, gt = 0;
      for (int tile = cta_id; tile < p.num_tiles; tile += cta_n, ++gt) {
        if (PAIR) mbar_wait_wd(&tempty_bar[acc], acc_ph ^ 1); else mbar_wait(&tempty_bar[acc], acc_ph ^ 1);     // epilogue has drained this accumulator
        if (lane == 0) ITN_TRACE_AT(7, gt);
        tc_fence_after();
        const uint32_t tacc = tmem_base + acc * Cfg::kAccCols;
        for (int kb = 0; kb < num_kb; ++kb, ++gk) {
          if (PAIR) mbar_wait_wd(&split_bar[s], ph); else mbar_wait(X3 ? &split_bar[s] : &full_bar[s], ph);
          tc_fence_after();
          if (elect_one()) {
            ITN_TRACE_AT(3, gk);
            const uint32_t sa = smem_u32(smem + s * Cfg::kStageBytes);
            const uint32_t sb = sa + Cfg::kABytes;
            const bool narrow = !A_MN && p.conv_kw > 0 && p.conv_c == 4;     // unswizzled tap slabs (see the producer)
#pragma unroll
            for (int k = 0; k < kBK / 8; ++k) {
              // K-major (SWIZZLE_128B): rows of 32 floats, 8-row groups 1024 B apart (SBO); the
              //   k-th MMA starts 8 floats = 32 B further along the swizzled row.
              // MN-major (SWIZZLE_128B_BASE32B): k-rows of 32 mn-floats, 4-row swizzle atoms 512 B
              //   apart (SBO), 32-wide mn blocks one TMA box = 4096 B apart (LBO); the k-th MMA
              //   starts 8 k-rows = 1024 B further.
              constexpr uint32_t kKLayout = kBK == 32 ? kLayoutSW128 : kLayoutSW64;   // K-major: rows of kBK floats
              constexpr uint32_t kKSbo = 8 * kBK * 4;                                 // 8-row groups
              const uint64_t ad = A_MN ? umma_smem_desc(sa + k * 1024, kAtomBytes, 512, kLayoutSW128Base32)
                                  : narrow ? umma_smem_desc(sa + k * (2 * kBM * 16), kBM * 16, 128, 0u)   // 2 tap slabs per MMA
                                           : umma_smem_desc(sa + k * 32, 16, kKSbo, kKLayout);
              const uint64_t bd = B_MN ? umma_smem_desc(sb + k * 1024, kAtomBytes, 512, kLayoutSW128Base32)
                                       : umma_smem_desc(sb + k * 32, 16, kKSbo, kKLayout);
              constexpr uint64_t kLoOff = static_cast<uint64_t>(Cfg::kRawBytes >> 4);
              if (PAIR) {                    // one 256 x BN x 8 product over both CTAs' tiles (same smem offsets)
                umma_tf32_pair(tacc, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
                umma_tf32_pair(tacc, ad + kLoOff, bd, idesc, 1u);
                umma_tf32_pair(tacc, ad, bd + kLoOff, idesc, 1u);
              } else {
                umma_tf32(tacc, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
                if (X3 && !(p.dbg & 32)) {     // dbg 32 (timing experiments only): main product alone
                  // residual tiles live kRawBytes after their raw twins, same layout: the descriptor
                  // start address is in 16-byte units
                  const bool split = Cfg::kCanSplit && p.split_acc;
                  const uint32_t tlo = split ? tacc + 2 * Cfg::kAccCols : tacc;
                  umma_tf32(tlo, ad + kLoOff, bd, idesc, (!split || (kb | k) != 0) ? 1u : 0u);   // A_lo * B_hi
                  umma_tf32(tlo, ad, bd + kLoOff, idesc, 1u);                                    // A_hi * B_lo
                }
              }
            }
            if (PAIR) {
              umma_commit_pair(&empty_bar[s]);                          // frees the slot in both CTAs
              if (kb == num_kb - 1) umma_commit_pair(&tfull_bar[acc]);  // both epilogues
            } else {
              umma_commit(&empty_bar[s]);  // frees the smem slot once the MMAs have read it
              ITN_TRACE_AT(4, gk);
              if (kb == num_kb - 1) umma_commit(&tfull_bar[acc]);  // accumulator complete
            }
          }
          __syncwarp();
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        acc ^= 1;
        if (acc == 0) acc_ph ^= 1;
      }
    }
  } else if (X3 && warp < Cfg::kEpiWarp0) {
    // ------------------------------------------------ residual splitters (tf32x3 mode)
    // lo = x - trunc_tf32(x), element-wise on the raw stage bytes (layout-agnostic, so the
    // swizzle is preserved), written kRawBytes further; then made visible to the async proxy.
    // (Pipelining the residual computation across k-blocks - pulling the next raw tile into
    // registers before the proxy fence of the current one - was measured and is slower: the fence
    // then also waits for those loads; 224 -> 160 TFLOP/s on 16480x2048x512.)
    int s = 0;
    uint32_t ph = 0;
    const int tid = threadIdx.x - 64;               // 0..127
    int gk = 0;
    for (int tile = cta_id; tile < p.num_tiles; tile += cta_n) {
      for (int kb = 0; kb < num_kb; ++kb, ++gk) {
        if (PAIR) mbar_wait_wd(&full_bar[s], ph); else mbar_wait(&full_bar[s], ph);
        if (tid == 0) ITN_TRACE_AT(1, gk);
        const float4* raw = reinterpret_cast<const float4*>(smem + s * Cfg::kStageBytes);
        float4* lo = reinterpret_cast<float4*>(smem + s * Cfg::kStageBytes + Cfg::kRawBytes);
        // Accumulator compensation.  Every tcgen05.mma adds into TMEM with round-toward-zero: an old
        // partial sum loses rz_eps (~0.6 * 2^-24) of its magnitude per later accumulate, i.e. the
        // k-block kb of a chain of 12*num_kb MMAs arrives weighted by 1 - rz_eps * (accumulates still
        // to come).  That weighting is linear in the contribution, so it is undone here, for free,
        // by folding  x * delta_kb  into the residual tile of A (delta_kb <= 5e-5 sits well inside
        // the 2^-11 range of x_lo).  Measured (tools/gemm_precision.py): K=2048 random operands
        // 1.5e-5 -> see profiles/README.md.
        const float kAcc = (Cfg::kCanSplit && p.split_acc ? 1.0f : 3.0f) * (kBK / 8);   // accumulates per k-block
        const float delta = p.rz_eps * (kAcc * static_cast<float>(num_kb - kb) - 0.5f * (kAcc - 1.0f)) + p.rz_c0;
        constexpr int kAVec = Cfg::kABytes / 16;
        if (!(p.dbg & 16))                 // dbg 16 (timing experiments only): no residual pass, garbage lo tiles
#pragma unroll 8
        for (int i = tid; i < (p.b_presplit && !B_MN ? Cfg::kABytes : Cfg::kRawBytes) / 16; i += 128) {
          const float4 x = raw[i];
          float4 r = tf32_residual(x);
          if (i < kAVec) {
            r.x = fmaf(x.x, delta, r.x); r.y = fmaf(x.y, delta, r.y);
            r.z = fmaf(x.z, delta, r.z); r.w = fmaf(x.w, delta, r.w);
          }
          lo[i] = r;
        }
        fence_proxy_async();
        __syncwarp();
        if (tid == 0) ITN_TRACE_AT(2, gk);
        if (lane == 0) { if (PAIR) mbar_arrive_cluster(&split_bar[s], 0); else mbar_arrive(&split_bar[s]); }
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp >= Cfg::kEpiWarp0) {
    // ----------------------------------------------------------- epilogue
    const int lg = warp & 3;    // TMEM lane group this warp may read: lanes [32*lg, 32*lg+32)
    float* st = staging + (warp - Cfg::kEpiWarp0) * (32 * 36);
    float4* st4 = reinterpret_cast<float4*>(st);
    const bool vec = p.vec != 0;
    int acc = 0;
    uint32_t acc_ph = 0;
    int gt = 0;
    for (int tile = cta_id; tile < p.num_tiles; tile += cta_n, ++gt) {
      const int nb = tile % p.tiles_n;
      const int t2 = tile / p.tiles_n;
      const int mb = PAIR ? 2 * (t2 % p.tiles_m) + (int)rank : t2 % p.tiles_m;
      const int z = t2 / p.tiles_m;
      const int n0 = nb * BN;
      const int row_base = mb * kBM + lg * 32;
      const EpiPtrs e = make_epi_ptrs(p, z / p.nb1, z % p.nb1);
      const int rmax = min(32, p.M - row_base);       // <= 0: this warp's rows are all padding
      constexpr int kChunks = BN / 32;
      const int nchunks = min(kChunks, (p.N - n0 + 31) / 32);   // live 32-column chunks (>= 1)
      // the streamed epilogue operand of the specialised variants (prefetched into registers)
      const int ev = p.variant;
      const float* xsp = nullptr;
      long long ldx = 0;
      if (ev == EV_RELU_MASK || ev == EV_GELU_GRAD) { xsp = e.aux; ldx = p.ldaux; }
      else if (ev == EV_RESIDUAL || ev == EV_RESIDUAL_RELU) { xsp = e.residual; ldx = p.ldr; }
      else if (ev == EV_ACCUMULATE) { xsp = e.C; ldx = p.ldc; }
      float4 xs4[8];
      auto prefetch = [&](int c) {     // 128-bit path only: 8 row-vectors of this lane's 4 columns
        const int col = n0 + c * 32 + (lane & 7) * 4;
        const bool ok = col < p.N;
        const float* src = xsp + (long long)(row_base + (lane >> 3)) * ldx + col;
#pragma unroll
        for (int i = 0; i < 8; ++i)
          xs4[i] = (ok && 4 * i + (lane >> 3) < rmax) ? *reinterpret_cast<const float4*>(src + (long long)(4 * i) * ldx)
                                                     : make_float4(0.f, 0.f, 0.f, 0.f);
      };
      // chunk 0 of the stream is fetched while the tensor core is still working on this tile
      if (vec && xsp && rmax > 0) prefetch(0);
      if (PAIR) mbar_wait_wd(&tfull_bar[acc], acc_ph); else mbar_wait(&tfull_bar[acc], acc_ph);
      if (warp == Cfg::kEpiWarp0 && lane == 0) ITN_TRACE_AT(5, gt);
      tc_fence_after();
      const uint32_t tacc = tmem_base + acc * Cfg::kAccCols + (static_cast<uint32_t>(lg * 32) << 16);
      if (rmax <= 0) {
        // nothing to store: hand the accumulator straight back to the issuer
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (PAIR) mbar_arrive_cluster(&tempty_bar[acc], 0); else mbar_arrive(&tempty_bar[acc]); }
      }
#pragma unroll 1
      for (int c = 0; c < nchunks && rmax > 0; ++c) {
        const int col0 = n0 + c * 32;
        if (vec && xsp && c > 0) prefetch(c);      // latency overlaps the TMEM read-out + transpose below
#ifdef ITN_TRACE
        const long long tp0 = clock64();
        long long tp1 = 0, tp2 = 0;
#endif
        {
          uint32_t v[32];
#ifdef ITN_TRACE
          if (p.dbg & 2) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = j + lane;
          } else
#endif
          {
            tmem_ld_32x32(tacc + c * 32, v);
            if (Cfg::kCanSplit && p.split_acc) {       // + the residual products' accumulator (fp32 RN add)
              uint32_t w[32];
              tmem_ld_32x32(tacc + 2 * Cfg::kAccCols + c * 32, w);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(w[j]));
            } else {
              tmem_ld_wait();
            }
          }
#ifdef ITN_TRACE
          tp1 = clock64();
          if (p.dbg & 4) {
            float sum = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) sum += __uint_as_float(v[j]);
            if (sum == 123.456f) st[0] = sum;
          } else
#endif
          if (vec) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              st4[lane * 9 + j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                              __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) st[lane * 33 + j] = __uint_as_float(v[j]);
          }
        }
        if (c == nchunks - 1) {
          // last chunk of this tile is out of TMEM: hand the accumulator back to the issuer
          tc_fence_before();
          __syncwarp();
          if (lane == 0) { if (PAIR) mbar_arrive_cluster(&tempty_bar[acc], 0); else mbar_arrive(&tempty_bar[acc]); }
        }
        __syncwarp();
#ifdef ITN_TRACE
        tp2 = clock64();
#endif
        if (vec) {
          const int col = col0 + (lane & 7) * 4;
#ifdef ITN_TRACE
          if (col < p.N && !(p.dbg & 1)) {
#else
          if (col < p.N) {
#endif
            switch (ev) {
              case EV_SIMPLE: epilogue_v4<EV_SIMPLE>(p, e, st4, lane, row_base, rmax, col, xs4); break;
              case EV_RESIDUAL: epilogue_v4<EV_RESIDUAL>(p, e, st4, lane, row_base, rmax, col, xs4); break;
              case EV_RESIDUAL_RELU: epilogue_v4<EV_RESIDUAL_RELU>(p, e, st4, lane, row_base, rmax, col, xs4); break;
              case EV_ACCUMULATE: epilogue_v4<EV_ACCUMULATE>(p, e, st4, lane, row_base, rmax, col, xs4); break;
              case EV_RELU_MASK: epilogue_v4<EV_RELU_MASK>(p, e, st4, lane, row_base, rmax, col, xs4); break;
              case EV_GELU_GRAD: epilogue_v4<EV_GELU_GRAD>(p, e, st4, lane, row_base, rmax, col, xs4); break;
              case EV_PRE: epilogue_v4<EV_PRE>(p, e, st4, lane, row_base, rmax, col, xs4); break;
              default: epilogue_v4<EV_PRE_GELU>(p, e, st4, lane, row_base, rmax, col, xs4); break;
            }
          }
        } else {
          const int col = col0 + lane;
          if (col < p.N) {
            switch (ev) {
              case EV_SIMPLE: epilogue_simple(p, e, st, lane, row_base, rmax, col); break;
              case EV_RESIDUAL: epilogue_rows<EV_RESIDUAL>(p, e, st, lane, row_base, rmax, col, xsp, ldx); break;
              case EV_RESIDUAL_RELU: epilogue_rows<EV_RESIDUAL_RELU>(p, e, st, lane, row_base, rmax, col, xsp, ldx); break;
              case EV_ACCUMULATE: epilogue_rows<EV_ACCUMULATE>(p, e, st, lane, row_base, rmax, col, xsp, ldx); break;
              case EV_RELU_MASK: epilogue_rows<EV_RELU_MASK>(p, e, st, lane, row_base, rmax, col, xsp, ldx); break;
              case EV_GELU_GRAD: epilogue_rows<EV_GELU_GRAD>(p, e, st, lane, row_base, rmax, col, xsp, ldx); break;
              case EV_PRE: epilogue_rows<EV_PRE>(p, e, st, lane, row_base, rmax, col, xsp, ldx); break;
              case EV_PRE_GELU: epilogue_rows<EV_PRE_GELU>(p, e, st, lane, row_base, rmax, col, xsp, ldx); break;
              default: epilogue_generic(p, e, st, lane, row_base, rmax, col); break;
            }
          }
        }
        __syncwarp();
#ifdef ITN_TRACE
        if (warp == Cfg::kEpiWarp0 && lane == 0 && blockIdx.x == 0 && g_trace != nullptr) {
          const long long tp3 = clock64();
          g_trace[8 * 1024 + 0] += tp1 - tp0;   // TMEM read
          g_trace[8 * 1024 + 1] += tp2 - tp1;   // smem transpose + hand-back
          g_trace[8 * 1024 + 2] += tp3 - tp2;   // math + global stores
          g_trace[8 * 1024 + 3] += 1;
        }
#endif
      }
      if (warp == Cfg::kEpiWarp0 && lane == 0) ITN_TRACE_AT(6, gt);
      acc ^= 1;
      if (acc == 0) acc_ph ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) ITN_TRACE_AT(8, 10);   // all roles done
  if (PAIR) cluster_sync_all();                // the peer may still be signalling our barriers / reading our tiles
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair<Cfg::kTmemCols>(tmem_base);
    else tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// ---------------------------------------------------------------- SIMT path
// CUDA-core fp32 GEMM with identical semantics, arbitrary alignment.
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const float* A, long long sam, long long sak, long long a_sb0, long long a_sb1,
                 const float* B, long long sbn, long long sbk, long long b_sb0, long long b_sb1,
                 const GemmKParams p) {
  pdl_wait();
  pdl_trigger();
  __shared__ float sA[16][33];
  __shared__ float sB[16][33];
  const int b0 = blockIdx.z / p.nb1, b1 = blockIdx.z % p.nb1;
  A += b0 * a_sb0 + b1 * a_sb1;
  B += b0 * b_sb0 + b1 * b_sb1;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  for (int k0 = 0; k0 < p.K; k0 += 16) {
    for (int i = threadIdx.x; i < 32 * 16; i += 256) {
      const int kk = i & 15, r = i >> 4;
      const int k = k0 + kk;
      const int m = m0 + r, n = n0 + r;
      sA[kk][r] = (m < p.M && k < p.K) ? A[m * sam + k * sak] : 0.f;
      sB[kk][r] = (n < p.N && k < p.K) ? B[n * sbn + k * sbk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const float a0 = sA[kk][ty], a1 = sA[kk][ty + 16];
      const float c0 = sB[kk][tx], c1 = sB[kk][tx + 16];
      acc[0][0] = fmaf(a0, c0, acc[0][0]);
      acc[0][1] = fmaf(a0, c1, acc[0][1]);
      acc[1][0] = fmaf(a1, c0, acc[1][0]);
      acc[1][1] = fmaf(a1, c1, acc[1][1]);
    }
    __syncthreads();
  }
  const EpiPtrs e = make_epi_ptrs(p, b0, b1);
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int row = m0 + ty + 16 * i, col = n0 + tx + 16 * j;
      if (row < p.M && col < p.N)
        epilogue_store(p, e, row, col, acc[i][j], e.bias ? e.bias[col] : 0.f);
    }
}

// ------------------------------------------------------------------- host
using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                              const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                              const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeFn get_encode_fn() {
  static EncodeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn>(f);
  });
  return fn;
}

static bool operand_tma_ok(const itn_operand_t& o, int nb0, int nb1) {
  if (reinterpret_cast<uintptr_t>(o.ptr) & 15) return false;
  if (o.ld <= 0 || (o.ld & 3)) return false;
  if (nb1 > 1 && o.sb1 != 0 && ((o.sb1 & 3) || o.sb1 < 0)) return false;
  if (nb0 > 1 && o.sb0 != 0 && ((o.sb0 & 3) || o.sb0 < 0)) return false;
  return true;
}

// rows = M (A) or N (B).  box_rows = tile extent along `rows` for the K-major case.
static int make_operand_map(CUtensorMap* tm, const itn_operand_t& o, int rows, int K, int nb0,
                            int nb1, int box_rows, int* mul0, int* mul1) {
  EncodeFn enc = get_encode_fn();
  if (!enc) return set_error(ITN_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  const bool bc0 = (nb0 <= 1) || o.sb0 == 0;
  const bool bc1 = (nb1 <= 1) || o.sb1 == 0;
  *mul0 = bc0 ? 0 : 1;
  *mul1 = bc1 ? 0 : 1;
  const cuuint64_t inner = o.major == 0 ? K : rows;
  const cuuint64_t outer = o.major == 0 ? rows : K;
  cuuint64_t gdim[4] = {inner, outer, bc1 ? 1ull : (cuuint64_t)nb1, bc0 ? 1ull : (cuuint64_t)nb0};
  const cuuint64_t dense = (cuuint64_t)o.ld * outer * 4;
  cuuint64_t gstr[3] = {(cuuint64_t)o.ld * 4, bc1 ? dense : (cuuint64_t)o.sb1 * 4,
                        bc0 ? dense : (cuuint64_t)o.sb0 * 4};
  cuuint32_t box[4] = {o.major == 0 ? (cuuint32_t)kBK : 32u, o.major == 0 ? (cuuint32_t)box_rows : (cuuint32_t)kBK, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(o.ptr), gdim, gstr,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   o.major == 0 ? (kBK == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B)
                                : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(ITN_ERR_CUDA,
                     "cuTensorMapEncodeTiled failed (%d): dims %llu %llu %llu %llu strides %llu "
                     "%llu %llu box %u %u",
                     (int)r, gdim[0], gdim[1], gdim[2], gdim[3], gstr[0], gstr[1], gstr[2], box[0],
                     box[1]);
  return ITN_OK;
}

using EncodeIm2colFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeIm2colFn get_encode_im2col_fn() {
  static EncodeIm2colFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeIm2colFn>(f);
  });
  return fn;
}

// Channels-last activation [n, h, w, c] as an im2col tensor map {c, w, h, n}: boxes of 32 channels x 128 output
// pixels; bounding-box corners -pad and pad - (k-1)*dil (forward convolution), traversal stride = conv stride.
static int make_im2col_map(CUtensorMap* tm, const itn_gemm_desc_t* d) {
  EncodeIm2colFn enc = get_encode_im2col_fn();
  if (!enc) return set_error(ITN_ERR_CUDA, "cuTensorMapEncodeIm2col entry point not available");
  cuuint64_t gdim[4] = {(cuuint64_t)d->conv_c, (cuuint64_t)d->conv_w, (cuuint64_t)d->conv_h, (cuuint64_t)d->conv_n};
  cuuint64_t gstr[3] = {(cuuint64_t)d->conv_c * 4, (cuuint64_t)d->conv_w * d->conv_c * 4,
                        (cuuint64_t)d->conv_h * d->conv_w * d->conv_c * 4};
  int lower[2] = {-d->conv_pad, -d->conv_pad};
  int upper[2] = {d->conv_pad - (d->conv_kw - 1) * d->conv_dil, d->conv_pad - (d->conv_kh - 1) * d->conv_dil};
  cuuint32_t estr[4] = {1, (cuuint32_t)d->conv_stride, (cuuint32_t)d->conv_stride, 1};
  // 4-channel input: boxes of [128 pixels x 4 channels = 16 bytes], unswizzled (one load per filter tap)
  const bool narrow = d->conv_c == 4;
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(d->A.ptr), gdim, gstr, lower, upper,
                   (cuuint32_t)(narrow ? 4 : kBK), (cuuint32_t)kBM, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   narrow ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(ITN_ERR_CUDA, "cuTensorMapEncodeIm2col failed (%d): [%d,%d,%d,%d] k %dx%d stride %d pad %d dil %d",
                     (int)r, d->conv_n, d->conv_h, d->conv_w, d->conv_c, d->conv_kh, d->conv_kw, d->conv_stride,
                     d->conv_pad, d->conv_dil);
  return ITN_OK;
}

static void fill_kparams(GemmKParams& p, const itn_gemm_desc_t* d) {
  p.M = d->M; p.N = d->N; p.K = d->K;
  p.conv_kw = d->conv_kh > 0 ? d->conv_kw : 0;
  p.conv_kh = d->conv_kh;
  p.conv_c = d->conv_c; p.conv_stride = d->conv_stride; p.conv_pad = d->conv_pad; p.conv_dil = d->conv_dil;
  p.conv_wo = d->conv_wo; p.conv_howo = d->conv_ho * d->conv_wo;
  p.nb1 = d->nb1 < 1 ? 1 : d->nb1;
  p.tiles_m = p.tiles_n = p.num_tiles = 0;
  p.a_m0 = p.a_m1 = p.b_m0 = p.b_m1 = 0;
  p.C = d->C; p.ldc = d->ldc; p.c_sb0 = d->c_sb0; p.c_sb1 = d->c_sb1;
  p.bias = d->bias; p.bias_sb0 = d->bias_sb0; p.bias_sb1 = d->bias_sb1;
  p.residual = d->residual; p.ldr = d->ldr; p.r_sb0 = d->r_sb0; p.r_sb1 = d->r_sb1;
  p.aux = d->aux; p.ldaux = d->ldaux; p.aux_sb0 = d->aux_sb0; p.aux_sb1 = d->aux_sb1;
  p.C2 = d->C2; p.ldc2 = d->ldc2; p.c2_sb0 = d->c2_sb0; p.c2_sb1 = d->c2_sb1;
  p.alpha = d->alpha; p.act = d->act; p.epi = d->epi; p.accumulate = d->accumulate;
  p.round_out = d->round_out;
  p.act_pos = d->act_pos;
  p.variant = epilogue_variant(d->aux != nullptr, d->residual != nullptr, d->C2 != nullptr, d->act, d->epi,
                               d->accumulate, d->act_pos);
  auto al = [](const void* ptr, long long ld, long long s0, long long s1) {
    return ptr == nullptr || (((reinterpret_cast<uintptr_t>(ptr) & 15) == 0) && (ld % 4 == 0) && (s0 % 4 == 0) &&
                              (s1 % 4 == 0));
  };
  // N % 4 != 0 can still take the 128-bit path when the caller owns the row padding (c_pad): the last
  // vector of a row then also covers up to 3 pad columns, whose accumulators are exact zeros (TMA
  // zero-fills the out-of-range B rows).  Plain bias-free epilogues only (no streamed operand to over-read).
  const bool pad_ok = d->c_pad && p.variant == EV_SIMPLE && d->bias == nullptr && d->ldc >= (d->N + 3) / 4 * 4;
  p.vec = (p.variant != EV_GENERIC) && (d->N % 4 == 0 || pad_ok) && al(d->C, d->ldc, d->c_sb0, d->c_sb1) &&
          al(d->bias, 4, d->bias_sb0, d->bias_sb1) && al(d->residual, d->ldr, d->r_sb0, d->r_sb1) &&
          al(d->aux, d->ldaux, d->aux_sb0, d->aux_sb1) && al(d->C2, d->ldc2, d->c2_sb0, d->c2_sb1);
  if (getenv("ITN_GEMM_NOVEC")) p.vec = 0;
  p.dbg = getenv("ITN_GEMM_DBG") ? atoi(getenv("ITN_GEMM_DBG")) : 0;
  // mean loss per round-toward-zero accumulate, in units of 2^-24 (ITN_GEMM_RZ_COMP=0 disables)
  // Fitted on the scale bias of the product against float64 (tools/gemm_bias_probe.py, zero-mean operands, K = 256 / 512 /
  // 2048 and the 24-accumulate chains of the fused attention): the shrink is  0.57 * (accumulates still to come) + 8.5
  // units - the constant part is the tensor core truncating the residual operands to 11 bits plus the dropped lo*lo term.
  static const float rz = getenv("ITN_GEMM_RZ_COMP") ? (float)atof(getenv("ITN_GEMM_RZ_COMP")) : 0.57f;
  static const float c0 = getenv("ITN_GEMM_RZ_C0") ? (float)atof(getenv("ITN_GEMM_RZ_C0")) : (rz > 0.0f ? 8.5f : 0.0f);
  p.rz_eps = rz * 5.9604645e-8f;
  p.rz_c0 = c0 * 5.9604645e-8f;
  p.split_acc = 0;
  p.b_presplit = (d->B_lo != nullptr && d->precision != ITN_PREC_TF32 && d->B.major == 0 &&
                  (reinterpret_cast<uintptr_t>(d->B_lo) & 15) == 0) ? 1 : 0;
}

static int validate(const itn_gemm_desc_t* d) {
  ITN_REQUIRE(d != nullptr, "gemm: null descriptor");
  ITN_REQUIRE(d->M > 0 && d->N > 0 && d->K > 0, "gemm: M,N,K must be positive (%d,%d,%d)", d->M,
              d->N, d->K);
  ITN_REQUIRE(d->A.ptr && d->B.ptr && d->C, "gemm: null A/B/C");
  ITN_REQUIRE(d->nb0 >= 1 && d->nb1 >= 1, "gemm: batch counts must be >= 1");
  ITN_REQUIRE((long long)d->nb0 * d->nb1 <= 65535, "gemm: batch %d x %d too large", d->nb0, d->nb1);
  ITN_REQUIRE(d->epi == ITN_EPI_NONE || d->aux != nullptr, "gemm: epi mode %d needs aux", d->epi);
  ITN_REQUIRE(d->A.major == 0 || d->A.major == 1, "gemm: bad A.major");
  ITN_REQUIRE(d->B.major == 0 || d->B.major == 1, "gemm: bad B.major");
  ITN_REQUIRE(d->precision == ITN_PREC_TF32X3 || d->precision == ITN_PREC_TF32 || d->precision == ITN_PREC_TF32X3_SPLIT,
              "gemm: bad precision %d", d->precision);
  if (d->conv_kh > 0) {
    ITN_REQUIRE(kBK == 32, "gemm: implicit convolution needs 32-float k-blocks");
    ITN_REQUIRE(d->conv_kw > 0 && d->conv_stride > 0 && d->conv_dil > 0 && d->conv_pad >= 0, "gemm: bad convolution geometry");
    ITN_REQUIRE(d->conv_c > 0 && (d->conv_c % 32 == 0 || d->conv_c == 4),
                "gemm: implicit convolution needs channels %% 32 == 0, or 4 (got %d)", d->conv_c);
    ITN_REQUIRE(d->conv_c != 4 || d->A.major == 0, "gemm: the 4-channel implicit convolution takes a K-major activation");
    ITN_REQUIRE(d->nb0 == 1 && d->nb1 == 1, "gemm: implicit convolution is not batched");
    ITN_REQUIRE((long long)d->conv_n * d->conv_ho * d->conv_wo == d->M && d->conv_kh * d->conv_kw * d->conv_c == d->K,
                "gemm: implicit convolution M/K do not match the geometry");
    ITN_REQUIRE(d->conv_ho == (d->conv_h + 2 * d->conv_pad - d->conv_dil * (d->conv_kh - 1) - 1) / d->conv_stride + 1 &&
                d->conv_wo == (d->conv_w + 2 * d->conv_pad - d->conv_dil * (d->conv_kw - 1) - 1) / d->conv_stride + 1,
                "gemm: implicit convolution output size does not match the geometry");
    ITN_REQUIRE((reinterpret_cast<uintptr_t>(d->A.ptr) & 15) == 0, "gemm: activation must be 16-byte aligned");
  }
  return ITN_OK;
}

// tf32x3 with the residual products in their own accumulator (GemmKParams::split_acc): asked for per call
// (ITN_PREC_TF32X3_SPLIT) or for every tf32x3 product of the process (ITN_GEMM_SPLITACC=1, experiments)
static bool want_split(const itn_gemm_desc_t* d) {
  static const int env = getenv("ITN_GEMM_SPLITACC") ? atoi(getenv("ITN_GEMM_SPLITACC")) : 0;
  return d->precision == ITN_PREC_TF32X3_SPLIT || (env && d->precision == ITN_PREC_TF32X3);
}

static int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

template <int BN, bool A_MN, bool B_MN, bool X3, bool PAIR = false>
static int launch_tile(const itn_gemm_desc_t* d, cudaStream_t stream) {
  using Cfg = TileCfg<BN, X3, PAIR>;
  GemmKParams p;
  fill_kparams(p, d);
  p.tiles_m = (d->M + kBM - 1) / kBM;
  if (PAIR) p.tiles_m = (p.tiles_m + 1) / 2;          // pairs of m-tiles: one 256-row tile per CTA pair
  p.tiles_n = (d->N + BN - 1) / BN;
  const long long nt = (long long)p.tiles_m * p.tiles_n * d->nb0 * d->nb1;
  if (nt > 0x7fffffffLL) return set_error(ITN_ERR_ARG, "gemm: too many tiles");
  p.num_tiles = (int)nt;
  p.split_acc = (X3 && Cfg::kCanSplit && want_split(d)) ? 1 : 0;
  if (p.split_acc && !getenv("ITN_GEMM_RZ_C0")) p.rz_c0 *= 5.0f / 8.5f;   // fitted (tools/gemm_bias_probe.py): the residual products no longer truncate the main sum
  CUtensorMap tmA, tmB, tmBlo;
  int rc = d->conv_kh > 0 ? make_im2col_map(&tmA, d)
                          : make_operand_map(&tmA, d->A, d->M, d->K, d->nb0, d->nb1, kBM, &p.a_m0, &p.a_m1);
  if (rc) return rc;
  rc = make_operand_map(&tmB, d->B, d->N, d->K, d->nb0, d->nb1, Cfg::kBRows, &p.b_m0, &p.b_m1);
  if (rc) return rc;
  tmBlo = tmB;
  if (p.b_presplit) {
    itn_operand_t blo = d->B;
    blo.ptr = d->B_lo;
    int m0 = 0, m1 = 0;
    rc = make_operand_map(&tmBlo, blo, d->N, d->K, d->nb0, d->nb1, Cfg::kBRows, &m0, &m1);
    if (rc) return rc;
  }
  auto kern = gemm_tf32_kernel<BN, A_MN, B_MN, X3, PAIR>;
  static bool attr_set = false;  // per instantiation
  if (!attr_set) {
    cudaError_t e =
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess)
      return set_error(ITN_ERR_CUDA, "gemm: cudaFuncSetAttribute(%d B): %s", Cfg::kSmemBytes,
                       cudaGetErrorString(e));
    attr_set = true;
  }
  if (PAIR) {
    const int pairs = sm_count() / 2;
    const int grid = 2 * (p.num_tiles < pairs ? p.num_tiles : pairs);   // persistent: one CTA pair per TPC
    launch_cluster(2, kern, grid, Cfg::kThreads, Cfg::kSmemBytes, stream, tmA, tmB, tmBlo, p);
    return check_launch("gemm_tf32_kernel (CTA pairs)");
  }
  int grid = p.num_tiles < sm_count() ? p.num_tiles : sm_count();   // persistent: <= 1 CTA per SM
  // A CTA walks tiles blockIdx.x, blockIdx.x + grid, ... with n fastest.  When the last n-tile is partial
  // (N=361 at BN=256: 8 live 32-column chunks, then 4) and grid shares a factor with tiles_n, every CTA sees
  // the same n positions over and over: half of them only full tiles, half only partial ones, and the
  // epilogue-bound products (K=32 attention scores) wait for the heavy half.  A grid coprime to tiles_n
  // rotates the positions (147 CTAs instead of 148: measured 325 -> see profiles/README.md).
  if (p.tiles_n > 1 && d->N % BN != 0 && grid == sm_count()) {
    auto gcd = [](int a, int b) { while (b) { const int t = a % b; a = b; b = t; } return a; };
    while (grid > 1 && gcd(grid, p.tiles_n) != 1) --grid;
  }
  launch(kern, grid, Cfg::kThreads, Cfg::kSmemBytes, stream, tmA, tmB, tmBlo, p);
  return check_launch("gemm_tf32_kernel");
}

// CTA pairs (cta_group::2) pay off where the 128 x 256 tf32x3 main loop is shared-memory bound: K-major operands,
// enough 256-row tiles to fill the 74 TPCs.  ITN_GEMM_PAIR=0/1 overrides (experiments).
static bool want_pair(const itn_gemm_desc_t* d) {
  static const int env = getenv("ITN_GEMM_PAIR") ? atoi(getenv("ITN_GEMM_PAIR")) : ITN_GEMM_PAIR_DEFAULT;
  if (!env || d->precision != ITN_PREC_TF32X3 || want_split(d) || d->A.major != 0 || d->B.major != 0 || kBK != 32) return false;
  const long long tiles_m = (d->M + kBM - 1) / kBM, tiles_n = (d->N + 255) / 256;
  const long long pair_tiles = ((tiles_m + 1) / 2) * tiles_n * d->nb0 * d->nb1;
  // measured (tools/gemm_pair_check.py): +12 % at K = 2048 (57760 x 256), +4 % at K = 1496, -1 ... -3 % at K <= 512
  // where the epilogue, not the main loop, sets the pace
  static const int kmin = getenv("ITN_GEMM_PAIR_KMIN") ? atoi(getenv("ITN_GEMM_PAIR_KMIN")) : 1024;
  return tiles_m >= 2 && d->N >= 192 && d->K >= kmin && pair_tiles >= sm_count() / 2;
}

template <int BN, bool X3>
static int launch_major2(const itn_gemm_desc_t* d, cudaStream_t s) {
  if constexpr (BN == 256 && X3 && kBK == 32) {
    if (want_pair(d)) return launch_tile<256, false, false, true, true>(d, s);
  }
  if (d->A.major == 0) {
    return d->B.major == 0 ? launch_tile<BN, false, false, X3>(d, s)
                           : launch_tile<BN, false, true, X3>(d, s);
  }
  return d->B.major == 0 ? launch_tile<BN, true, false, X3>(d, s)
                         : launch_tile<BN, true, true, X3>(d, s);
}

template <int BN>
static int launch_major(const itn_gemm_desc_t* d, cudaStream_t s) {
  return d->precision == ITN_PREC_TF32 ? launch_major2<BN, false>(d, s) : launch_major2<BN, true>(d, s);   // X3: either tf32x3 flavour
}

// Tile width: the widest tile that still yields at least one tile per SM (wide tiles re-read
// less from L2: 48 KB of operands per 128x256x32 step vs 32 KB per 128x128x32); small problems
// fall back to 64-wide tiles to spread over more SMs.
static int pick_bn(const itn_gemm_desc_t* d) {
  if (d->N <= 32) return 32;
  if (d->N <= 64) return 64;
  const long long tiles_m = (d->M + kBM - 1) / kBM;
  const long long batch = (long long)d->nb0 * d->nb1;
  const int cand[3] = {256, 128, 64};
  for (int i = 0; i < 3; ++i) {
    const int bn = cand[i];
    if (bn > 64 && d->N <= bn / 2) continue;
    const long long tiles = tiles_m * ((d->N + bn - 1) / bn) * batch;
    if (tiles >= sm_count()) return bn;
  }
  return 64;
}

}  // namespace itn

#ifdef ITN_TRACE
extern "C" int itn_debug_set_trace(long long* buf) {
  return cudaMemcpyToSymbol(itn::g_trace, &buf, sizeof(buf)) == cudaSuccess ? 0 : -2;
}
#endif

extern "C" int itn_gemm_tf32_supported(const itn_gemm_desc_t* d) {
  if (!d || d->M <= 0 || d->N <= 0 || d->K <= 0) return 0;
  const bool a_ok = d->conv_kh > 0 ? ((d->conv_c % 32 == 0 || d->conv_c == 4) && (reinterpret_cast<uintptr_t>(d->A.ptr) & 15) == 0)
                                   : itn::operand_tma_ok(d->A, d->nb0, d->nb1);
  return a_ok && itn::operand_tma_ok(d->B, d->nb0, d->nb1);
}

extern "C" int itn_gemm_tf32(const itn_gemm_desc_t* d, void* stream) {
  int rc = itn::validate(d);
  if (rc) return rc;
  if (!itn_gemm_tf32_supported(d))
    return itn::set_error(ITN_ERR_UNSUPPORTED,
                          "gemm_tf32: operands must be 16-byte aligned with ld/batch strides "
                          "multiples of 4 elements");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int bn = itn::pick_bn(d);
  if (bn > 128 && itn::want_split(d)) bn = 128;     // the split accumulators need half of tensor memory each
  if (const char* f = getenv("ITN_GEMM_BN")) {
    const int v = atoi(f);
    if (v == 32 || v == 64 || v == 128 || v == 256) bn = v;
  }
  switch (bn) {
    case 32: return itn::launch_major<32>(d, s);
    case 64: return itn::launch_major<64>(d, s);
    case 128: return itn::launch_major<128>(d, s);
    default: return itn::launch_major<256>(d, s);
  }
}

extern "C" int itn_gemm_simt(const itn_gemm_desc_t* d, void* stream) {
  int rc = itn::validate(d);
  if (rc) return rc;
  itn::GemmKParams p;
  itn::fill_kparams(p, d);
  const long long sam = d->A.major == 0 ? d->A.ld : 1, sak = d->A.major == 0 ? 1 : d->A.ld;
  const long long sbn = d->B.major == 0 ? d->B.ld : 1, sbk = d->B.major == 0 ? 1 : d->B.ld;
  dim3 grid((d->N + 31) / 32, (d->M + 31) / 32, d->nb0 * d->nb1);
  itn::launch(itn::gemm_simt_kernel, grid, 256, 0, static_cast<cudaStream_t>(stream), 
      d->A.ptr, sam, sak, d->A.sb0, d->A.sb1, d->B.ptr, sbn, sbk, d->B.sb0, d->B.sb1, p);
  return itn::check_launch("gemm_simt_kernel");
}
