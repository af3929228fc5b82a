// Trainer step over the flat parameter / meta-gradient buffers:
//   torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm)     engine/interactron_trainer.py:106
//   detector_optimizer.step(); supervisor_optimizer.step()            :107-108   (torch.optim.Adam)
//   *.zero_grad()                                                     :109-110
// as two HBM-bound kernels: a fixed-order sum of squares of the gradient buffer (one partial per
// CTA) and one pass that finishes the norm (every CTA re-reduces the <= 1184 partials in the same
// order, so the clip coefficient is identical everywhere and run-to-run), scales the gradient,
// updates both Adam moments and the weights, and optionally zeroes the gradient.
// Bytes per element: 4 (g) in the first pass; 16 read (g, w, m, v) + 12-16 written in the second.
#include "itn_common.cuh"

namespace itn {

constexpr int kSMs = 148;
constexpr int kNormThreads = 256;
constexpr int kMaxPartials = kSMs * 8;
// "this parameter has no gradient" (torch: .grad is None -> clip_grad_norm_ and Adam skip it): a quiet
// NaN with a payload no arithmetic produces.  Compared bitwise, so genuine NaNs still propagate.
constexpr unsigned kNoGradBits = ITN_NO_GRAD_BITS;
__device__ __forceinline__ bool no_grad(float g) { return __float_as_uint(g) == kNoGradBits; }
__device__ __forceinline__ float sq_or_0(float g) { return no_grad(g) ? 0.f : g * g; }

__device__ __forceinline__ float block_sum_256(float v, float* sh) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x < 32) {
    t = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) sh[0] = t;
  }
  __syncthreads();
  t = sh[0];
  __syncthreads();
  return t;
}

// partial[blockIdx.x] = sum of g[i]^2 over this CTA's grid-stride slice (fixed assignment).
__global__ void __launch_bounds__(kNormThreads)
sumsq_partial_kernel(const float* __restrict__ g, long long n, float* __restrict__ partial) {
  pdl_wait();
  pdl_trigger();
  __shared__ float sh[32];
  const long long n4 = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  const long long stride = (long long)gridDim.x * blockDim.x;
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = g4[i];
    acc += sq_or_0(v.x); acc += sq_or_0(v.y);
    acc += sq_or_0(v.z); acc += sq_or_0(v.w);
  }
  if (blockIdx.x == 0)
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) acc += sq_or_0(g[i]);
  const float t = block_sum_256(acc, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

struct AdamArgs {
  float step_size;  // lr / (1 - beta1^step)
  float beta2, eps;
  float omb1, omb2; // 1 - beta1, 1 - beta2 rounded from double (the Python scalars torch passes)
  float bc2_sqrt;   // sqrt(1 - beta2^step)
  float max_norm;   // <= 0: no clipping
  int zero_grad;
};

// torch.optim.Adam._single_tensor_adam (weight_decay 0, amsgrad off), same operation order:
//   m.lerp_(g, 1-b1); v.mul_(b2).addcmul_(g, g, 1-b2); denom = sqrt(v)/sqrt(bc2) + eps;
//   w.addcdiv_(m, denom, value=-(lr/bc1))
__device__ __forceinline__ void adam_one(float& w, float g, float& m, float& v, const AdamArgs& a,
                                         float coef, float step_size) {
  if (no_grad(g)) return;
  g *= coef;
  m = fmaf(g - m, a.omb1, m);
  v = fmaf(g * g, a.omb2, v * a.beta2);
  const float denom = sqrtf(v) / a.bc2_sqrt + a.eps;
  w = fmaf(-step_size, m / denom, w);
}

__global__ void __launch_bounds__(256)
clip_adam_kernel(float* __restrict__ w, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                 long long n, const float* __restrict__ partial, int n_partial, float* __restrict__ norm_out,
                 const AdamArgs a) {
  pdl_wait();
  pdl_trigger();
  __shared__ float sh[32];
  float coef = 1.0f;
  if (partial) {
    float acc = 0.f;
    for (int i = threadIdx.x; i < n_partial; i += blockDim.x) acc += partial[i];
    const float total = sqrtf(block_sum_256(acc, sh));
    if (norm_out && blockIdx.x == 0 && threadIdx.x == 0) *norm_out = total;
    if (a.max_norm > 0.f) coef = fminf(a.max_norm / (total + 1e-6f), 1.0f);   // clip_grad_norm_: clamp(max=1)
  }
  const float step_size = a.step_size;
  const long long n4 = n >> 2;
  float4* w4 = reinterpret_cast<float4*>(w);
  float4* g4 = reinterpret_cast<float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 ww = w4[i], mm = m4[i], vv = v4[i];
    const float4 gg = g4[i];
    adam_one(ww.x, gg.x, mm.x, vv.x, a, coef, step_size);
    adam_one(ww.y, gg.y, mm.y, vv.y, a, coef, step_size);
    adam_one(ww.z, gg.z, mm.z, vv.z, a, coef, step_size);
    adam_one(ww.w, gg.w, mm.w, vv.w, a, coef, step_size);
    w4[i] = ww; m4[i] = mm; v4[i] = vv;
    if (a.zero_grad) g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (blockIdx.x == 0) {
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) {
      float ww = w[i], mm = m[i], vv = v[i];
      adam_one(ww, g[i], mm, vv, a, coef, step_size);
      w[i] = ww; m[i] = mm; v[i] = vv;
      if (a.zero_grad) g[i] = 0.f;
    }
  }
}


// acc = (first ? 0 : acc) + w * x, in torch's operation order (product rounded, then the sum rounded: no FMA
// contraction), float4 grid-stride.  The windowed checkpoint average of the trainer.
__global__ void __launch_bounds__(256) ckpt_accumulate_kernel(const float* __restrict__ x, float* __restrict__ acc,
                                                              long long n, float w, int first) {
  pdl_wait();
  pdl_trigger();
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    float4 a = first ? make_float4(0.f, 0.f, 0.f, 0.f) : reinterpret_cast<float4*>(acc)[i];
    if (first) {
      a = make_float4(__fmul_rn(w, v.x), __fmul_rn(w, v.y), __fmul_rn(w, v.z), __fmul_rn(w, v.w));
    } else {
      a.x = __fadd_rn(a.x, __fmul_rn(w, v.x));
      a.y = __fadd_rn(a.y, __fmul_rn(w, v.y));
      a.z = __fadd_rn(a.z, __fmul_rn(w, v.z));
      a.w = __fadd_rn(a.w, __fmul_rn(w, v.w));
    }
    reinterpret_cast<float4*>(acc)[i] = a;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (long long i = n4 << 2; i < n; ++i) acc[i] = first ? __fmul_rn(w, x[i]) : __fadd_rn(acc[i], __fmul_rn(w, x[i]));
}

}  // namespace itn

using namespace itn;

extern "C" int itn_sumsq_partials(const float* g, long long n, float* partial, int max_partials,
                                  int* n_partials, void* stream) {
  ITN_REQUIRE(g && partial && n_partials && n > 0 && max_partials > 0, "sumsq_partials: bad arguments");
  ITN_REQUIRE(((uintptr_t)g & 15) == 0, "sumsq_partials: g must be 16-byte aligned");
  long long blocks = ((n >> 2) + kNormThreads * 8 - 1) / (kNormThreads * 8);   // >= 8 float4 per thread
  if (blocks > kMaxPartials) blocks = kMaxPartials;
  if (blocks > max_partials) blocks = max_partials;
  if (blocks < 1) blocks = 1;
  *n_partials = (int)blocks;
  launch(sumsq_partial_kernel, dim3((unsigned)blocks), kNormThreads, 0, static_cast<cudaStream_t>(stream), g, n,
         partial);
  return check_launch("sumsq_partial_kernel");
}

extern "C" int itn_clip_adam_step(float* w, float* g, float* m, float* v, long long n, const float* partial,
                                  int n_partials, float max_norm, double lr, double beta1, double beta2,
                                  double eps, int step, int zero_grad, float* norm_out, void* stream) {
  ITN_REQUIRE(w && g && m && v && n > 0 && step >= 1, "clip_adam_step: bad arguments");
  ITN_REQUIRE(((((uintptr_t)w | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v)) & 15) == 0,
              "clip_adam_step: buffers must be 16-byte aligned");
  ITN_REQUIRE(!partial || (n_partials > 0 && n_partials <= kMaxPartials), "clip_adam_step: bad partial count");
  AdamArgs a;
  // scalars are derived in double, as the Python floats of torch.optim.Adam are, and rounded once
  const double bc1 = 1.0 - pow(beta1, (double)step);
  a.step_size = (float)(lr / bc1);
  a.beta2 = (float)beta2; a.eps = (float)eps;
  a.omb1 = (float)(1.0 - beta1);
  a.omb2 = (float)(1.0 - beta2);
  a.bc2_sqrt = (float)sqrt(1.0 - pow(beta2, (double)step));
  a.max_norm = max_norm;
  a.zero_grad = zero_grad;
  long long blocks = ((n >> 2) + 255) / 256;
  if (blocks > (long long)kSMs * 8) blocks = (long long)kSMs * 8;
  if (blocks < 1) blocks = 1;
  launch(clip_adam_kernel, dim3((unsigned)blocks), 256, 0, static_cast<cudaStream_t>(stream), w, g, m, v, n,
         partial, n_partials, norm_out, a);
  return check_launch("clip_adam_kernel");
}

extern "C" int itn_ckpt_accumulate(const float* x, float* acc, long long n, float w, int first, void* stream) {
  ITN_REQUIRE(x && acc && n > 0, "ckpt_accumulate: bad arguments");
  ITN_REQUIRE((((uintptr_t)x | (uintptr_t)acc) & 15) == 0, "ckpt_accumulate: pointers must be 16-byte aligned");
  long long blocks = ((n >> 2) + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 8) blocks = 148 * 8;
  itn::launch(itn::ckpt_accumulate_kernel, (int)blocks, 256, 0, static_cast<cudaStream_t>(stream), x, acc, n, w, first);
  return itn::check_launch("ckpt_accumulate_kernel");
}
