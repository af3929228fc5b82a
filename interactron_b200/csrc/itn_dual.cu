// Tangent (forward-mode) companions of the row-wise / element-wise kernels, used by the
// dual-number pass that yields the second-order meta-gradient of the training step
// (interactron_b200/dual.py; reference models/interactron.py:98-123 obtains the same numbers by a
// double backward).  Same layout rules as itn_rowwise.cu: one warp per row, float4 per lane,
// warp-shuffle reductions, every operand read once.  Any "*_dot" pointer may be null = zero.
#include "itn_common.cuh"

namespace itn {

__device__ __forceinline__ float4 ld4(const float* p, long long i) {
  return reinterpret_cast<const float4*>(p)[i];
}
__device__ __forceinline__ float4 ld4z(const float* p, long long i) {
  return p ? reinterpret_cast<const float4*>(p)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
}

// ------------------------------------------------------------------ LayerNorm forward tangent
// y_dot = gamma * xhat_dot + gamma_dot * xhat + beta_dot,
// xhat_dot = rstd * (xc - xhat * mean(xhat * xc)),  xc = x_dot - mean(x_dot).
template <int VPL>
__global__ void __launch_bounds__(256)
layernorm_fwd_jvp_kernel(const float* __restrict__ x, const float* __restrict__ x_dot,
                         const float* __restrict__ mean, const float* __restrict__ rstd,
                         const float* __restrict__ gamma, const float* __restrict__ gamma_dot,
                         const float* __restrict__ beta_dot, float* __restrict__ y_dot, long long rows,
                         long long rpg_g, long long g_stride, long long rpg_d, long long d_stride) {
  pdl_wait();
  pdl_trigger();
  constexpr int cols = VPL * 128;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float mu = mean[row], rs = rstd[row];
  const float* xr = x + row * cols;
  const float* xd = x_dot ? x_dot + row * cols : nullptr;
  float4 xh[VPL], xc[VPL];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const float4 xv = ld4(xr, lane + 32 * i);
    xc[i] = ld4z(xd, lane + 32 * i);
    xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
    s += (xc[i].x + xc[i].y) + (xc[i].z + xc[i].w);
  }
  const float md = warp_sum(s) * (1.0f / cols);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    xc[i].x -= md; xc[i].y -= md; xc[i].z -= md; xc[i].w -= md;
    q += (xh[i].x * xc[i].x + xh[i].y * xc[i].y) + (xh[i].z * xc[i].z + xh[i].w * xc[i].w);
  }
  const float m = warp_sum(q) * (1.0f / cols);
  const float* gr = gamma + (row / rpg_g) * g_stride;
  const float* gd = gamma_dot ? gamma_dot + (row / rpg_d) * d_stride : nullptr;
  const float* bd = beta_dot ? beta_dot + (row / rpg_d) * d_stride : nullptr;
  float4* yr = reinterpret_cast<float4*>(y_dot + row * cols);
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const float4 ga = ld4(gr, lane + 32 * i), gdv = ld4z(gd, lane + 32 * i), bdv = ld4z(bd, lane + 32 * i);
    float4 o;
    o.x = ga.x * rs * (xc[i].x - xh[i].x * m) + gdv.x * xh[i].x + bdv.x;
    o.y = ga.y * rs * (xc[i].y - xh[i].y * m) + gdv.y * xh[i].y + bdv.y;
    o.z = ga.z * rs * (xc[i].z - xh[i].z * m) + gdv.z * xh[i].z + bdv.z;
    o.w = ga.w * rs * (xc[i].w - xh[i].w * m) + gdv.w * xh[i].w + bdv.w;
    yr[lane + 32 * i] = o;
  }
}

// ------------------------------------------------------------------ LayerNorm backward tangent
// primal: u = gamma*dy, a = mean(u), b = mean(u*xhat), dx = rstd*(u - a - xhat*b).
// tangent: u_dot = gamma_dot*dy + gamma*dy_dot, a_dot = mean(u_dot), b_dot = mean(u_dot*xhat + u*xhat_dot),
//          dx_dot = -rstd*m*dx + rstd*(u_dot - a_dot - xhat_dot*b - xhat*b_dot)   (rstd_dot = -rstd^2 * m).
// gterm (optional) receives dy_dot*xhat + dy*xhat_dot, whose column sums are the tangent of dgamma.
template <int VPL>
__global__ void __launch_bounds__(256)
layernorm_bwd_jvp_kernel(const float* __restrict__ dy, const float* __restrict__ dy_dot,
                         const float* __restrict__ x, const float* __restrict__ x_dot,
                         const float* __restrict__ mean, const float* __restrict__ rstd,
                         const float* __restrict__ gamma, const float* __restrict__ gamma_dot,
                         float* __restrict__ dx_dot, float* __restrict__ gterm, long long rows, long long rpg_g,
                         long long g_stride, long long rpg_d, long long d_stride) {
  pdl_wait();
  pdl_trigger();
  constexpr int cols = VPL * 128;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float mu = mean[row], rs = rstd[row];
  const float* xr = x + row * cols;
  const float* xd = x_dot ? x_dot + row * cols : nullptr;
  const float* dr = dy + row * cols;
  const float* dd = dy_dot ? dy_dot + row * cols : nullptr;
  const float* gr = gamma + (row / rpg_g) * g_stride;
  const float* gd = gamma_dot ? gamma_dot + (row / rpg_d) * d_stride : nullptr;
  float4 xh[VPL], xc[VPL], u[VPL], ud[VPL];
  float s0 = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const float4 xv = ld4(xr, lane + 32 * i);
    xc[i] = ld4z(xd, lane + 32 * i);
    xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
    s0 += (xc[i].x + xc[i].y) + (xc[i].z + xc[i].w);
  }
  const float md = warp_sum(s0) * (1.0f / cols);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    xc[i].x -= md; xc[i].y -= md; xc[i].z -= md; xc[i].w -= md;
    q += (xh[i].x * xc[i].x + xh[i].y * xc[i].y) + (xh[i].z * xc[i].z + xh[i].w * xc[i].w);
  }
  const float m = warp_sum(q) * (1.0f / cols);
  // xc <- xhat_dot
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    xc[i].x = rs * (xc[i].x - xh[i].x * m); xc[i].y = rs * (xc[i].y - xh[i].y * m);
    xc[i].z = rs * (xc[i].z - xh[i].z * m); xc[i].w = rs * (xc[i].w - xh[i].w * m);
  }
  float sa = 0.f, sb = 0.f, sad = 0.f, sbd = 0.f;
  float4* gt = gterm ? reinterpret_cast<float4*>(gterm + row * cols) : nullptr;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const float4 dv = ld4(dr, lane + 32 * i), ddv = ld4z(dd, lane + 32 * i);
    const float4 ga = ld4(gr, lane + 32 * i), gdv = ld4z(gd, lane + 32 * i);
    u[i] = make_float4(ga.x * dv.x, ga.y * dv.y, ga.z * dv.z, ga.w * dv.w);
    ud[i] = make_float4(gdv.x * dv.x + ga.x * ddv.x, gdv.y * dv.y + ga.y * ddv.y, gdv.z * dv.z + ga.z * ddv.z,
                        gdv.w * dv.w + ga.w * ddv.w);
    sa += (u[i].x + u[i].y) + (u[i].z + u[i].w);
    sb += (u[i].x * xh[i].x + u[i].y * xh[i].y) + (u[i].z * xh[i].z + u[i].w * xh[i].w);
    sad += (ud[i].x + ud[i].y) + (ud[i].z + ud[i].w);
    sbd += (ud[i].x * xh[i].x + u[i].x * xc[i].x) + (ud[i].y * xh[i].y + u[i].y * xc[i].y) +
           (ud[i].z * xh[i].z + u[i].z * xc[i].z) + (ud[i].w * xh[i].w + u[i].w * xc[i].w);
    if (gt)
      gt[lane + 32 * i] = make_float4(ddv.x * xh[i].x + dv.x * xc[i].x, ddv.y * xh[i].y + dv.y * xc[i].y,
                                      ddv.z * xh[i].z + dv.z * xc[i].z, ddv.w * xh[i].w + dv.w * xc[i].w);
  }
  const float a = warp_sum(sa) * (1.0f / cols), b = warp_sum(sb) * (1.0f / cols);
  const float ad = warp_sum(sad) * (1.0f / cols), bd = warp_sum(sbd) * (1.0f / cols);
  float4* o4 = reinterpret_cast<float4*>(dx_dot + row * cols);
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    float4 o;
#define ITN_LNJ(c)                                              \
  {                                                             \
    const float dxp = rs * (u[i].c - a - xh[i].c * b);          \
    o.c = -rs * m * dxp + rs * (ud[i].c - ad - xc[i].c * b - xh[i].c * bd); \
  }
    ITN_LNJ(x) ITN_LNJ(y) ITN_LNJ(z) ITN_LNJ(w)
#undef ITN_LNJ
    o4[lane + 32 * i] = o;
  }
}

// ------------------------------------------------------------------ softmax backward, dual
// dp     <- dS     = scale * p * (dp - r),                         r  = sum p*dp
// dp_dot <- dS_dot = scale * [p_dot*(dp - r) + p*(dp_dot - rd)],   rd = sum (p_dot*dp + p*dp_dot)
template <int NV>
__global__ void __launch_bounds__(256)
softmax_bwd_jvp_kernel(const float* __restrict__ p, const float* __restrict__ p_dot, float* __restrict__ dp,
                       float* __restrict__ dp_dot, long long rows, int cols, long long ld, float scale) {
  pdl_wait();
  pdl_trigger();
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* pr = p + row * ld;
  const float* pdr = p_dot ? p_dot + row * ld : nullptr;
  float4* d4 = reinterpret_cast<float4*>(dp + row * ld);
  float4* dd4 = reinterpret_cast<float4*>(dp_dot + row * ld);
  const int nvec = (cols + 3) >> 2;
  float pv[NV][4], pd[NV][4], dv[NV][4], dd[NV][4];
  float r = 0.f, rd = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int j = lane + 32 * i;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a, c = a, e = a;
    if (j < nvec) {
      a = ld4(pr, j);
      b = ld4z(pdr, j);
      c = d4[j];
      e = dd4[j];
    }
    pv[i][0] = a.x; pv[i][1] = a.y; pv[i][2] = a.z; pv[i][3] = a.w;
    pd[i][0] = b.x; pd[i][1] = b.y; pd[i][2] = b.z; pd[i][3] = b.w;
    dv[i][0] = c.x; dv[i][1] = c.y; dv[i][2] = c.z; dv[i][3] = c.w;
    dd[i][0] = e.x; dd[i][1] = e.y; dd[i][2] = e.z; dd[i][3] = e.w;
    const int c0 = 4 * j;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (c0 + k >= cols) pv[i][k] = 0.f, pd[i][k] = 0.f, dv[i][k] = 0.f, dd[i][k] = 0.f;
      r = fmaf(pv[i][k], dv[i][k], r);
      rd = fmaf(pd[i][k], dv[i][k], fmaf(pv[i][k], dd[i][k], rd));
    }
  }
  r = warp_sum(r);
  rd = warp_sum(rd);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int j = lane + 32 * i;
    if (j < nvec) {
      float o[4], od[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        o[k] = scale * pv[i][k] * (dv[i][k] - r);
        od[k] = scale * (pd[i][k] * (dv[i][k] - r) + pv[i][k] * (dd[i][k] - rd));
      }
      d4[j] = make_float4(o[0], o[1], o[2], o[3]);
      dd4[j] = make_float4(od[0], od[1], od[2], od[3]);
    }
  }
}

// ------------------------------------------------------------------ element-wise
__global__ void __launch_bounds__(256)
mask_mul_kernel(float* __restrict__ y, const float* __restrict__ ref, long long n4) {
  pdl_wait();
  pdl_trigger();
  float4* y4 = reinterpret_cast<float4*>(y);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 v = y4[i];
    const float4 r = ld4(ref, i);
    v.x = r.x > 0.f ? v.x : 0.f; v.y = r.y > 0.f ? v.y : 0.f;
    v.z = r.z > 0.f ? v.z : 0.f; v.w = r.w > 0.f ? v.w : 0.f;
    y4[i] = v;
  }
}

__global__ void __launch_bounds__(256)
mul_mask_u8_kernel(const float* __restrict__ x, const unsigned char* __restrict__ mask, float scale,
                   float* __restrict__ out, long long n4) {
  pdl_wait();
  pdl_trigger();
  const uchar4* m4 = reinterpret_cast<const uchar4*>(mask);
  float4* o4 = reinterpret_cast<float4*>(out);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = ld4(x, i);
    const uchar4 m = m4[i];
    o4[i] = make_float4(m.x ? v.x * scale : 0.f, m.y ? v.y * scale : 0.f, m.z ? v.z * scale : 0.f,
                        m.w ? v.w * scale : 0.f);
  }
}

__device__ __forceinline__ void gelu_d12(float x, float& g1, float& g2) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
  g1 = cdf + x * pdf;
  g2 = pdf * (2.0f - x * x);
}

// y = raw*gelu'(aux);  y_dot = raw_dot*gelu'(aux) + raw*gelu''(aux)*aux_dot
__global__ void __launch_bounds__(256)
gelu_grad_dual_kernel(const float* __restrict__ raw, const float* __restrict__ raw_dot,
                      const float* __restrict__ aux, const float* __restrict__ aux_dot, float* __restrict__ y,
                      float* __restrict__ y_dot, long long n4) {
  pdl_wait();
  pdl_trigger();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 r = ld4(raw, i), rd = ld4z(raw_dot, i), a = ld4(aux, i), ad = ld4z(aux_dot, i);
    float4 o, od;
    float g1, g2;
    gelu_d12(a.x, g1, g2); o.x = r.x * g1; od.x = rd.x * g1 + r.x * g2 * ad.x;
    gelu_d12(a.y, g1, g2); o.y = r.y * g1; od.y = rd.y * g1 + r.y * g2 * ad.y;
    gelu_d12(a.z, g1, g2); o.z = r.z * g1; od.z = rd.z * g1 + r.z * g2 * ad.z;
    gelu_d12(a.w, g1, g2); o.w = r.w * g1; od.w = rd.w * g1 + r.w * g2 * ad.w;
    reinterpret_cast<float4*>(y)[i] = o;
    if (y_dot) reinterpret_cast<float4*>(y_dot)[i] = od;
  }
}

// dx_dot = dy_dot*y*(1-y) + dy*(1-2y)*y_dot   (tangent of dx = dy*y*(1-y))
__global__ void __launch_bounds__(256)
sigmoid_bwd_jvp_kernel(const float* __restrict__ dy, const float* __restrict__ dy_dot, const float* __restrict__ y,
                       const float* __restrict__ y_dot, float* __restrict__ out, long long n) {
  pdl_wait();
  pdl_trigger();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float yv = y[i];
    float o = 0.f;
    if (dy_dot) o += dy_dot[i] * yv * (1.0f - yv);
    if (y_dot) o += dy[i] * (1.0f - 2.0f * yv) * y_dot[i];
    out[i] = o;
  }
}

// n_dot = <d, x_dot>;  d_dot = (x_dot - d*n_dot)/nrm   (d = x/||x||); one block per group
__global__ void __launch_bounds__(256)
l2norm_jvp_kernel(const float* __restrict__ x_dot, const float* __restrict__ nrm, const float* __restrict__ d,
                  float* __restrict__ n_dot, float* __restrict__ d_dot, int n) {
  pdl_wait();
  pdl_trigger();
  __shared__ float part[8];
  const int g = blockIdx.x;
  const float* xd = x_dot + (long long)g * n;
  const float* dg = d + (long long)g * n;
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) s = fmaf(dg[i], xd[i], s);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) t += part[i];
  if (threadIdx.x == 0) n_dot[g] = t;
  const float inv = 1.0f / nrm[g];
  for (int i = threadIdx.x; i < n; i += 256) d_dot[(long long)g * n + i] = (xd[i] - dg[i] * t) * inv;
}

static inline unsigned ew_grid(long long n) {
  long long b = (n + 255) / 256;
  const long long cap = 148LL * 16;
  return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}
static inline bool al16(const void* p) { return (((uintptr_t)p) & 15) == 0; }

}  // namespace itn

using namespace itn;

extern "C" int itn_layernorm_fwd_jvp(const float* x, const float* x_dot, const float* mean, const float* rstd,
                                     const float* gamma, const float* gamma_dot, const float* beta_dot,
                                     float* y_dot, long long rows, int cols, int groups, long long g_stride,
                                     int groups_dot, long long d_stride, void* stream) {
  ITN_REQUIRE(x && mean && rstd && gamma && y_dot, "layernorm_fwd_jvp: null pointer");
  ITN_REQUIRE(rows > 0 && groups > 0 && groups_dot > 0 && rows % groups == 0 && rows % groups_dot == 0,
              "layernorm_fwd_jvp: rows (%lld) must be a multiple of groups (%d, %d)", rows, groups, groups_dot);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const unsigned grid = (unsigned)((rows + 7) / 8);
  const long long rg = rows / groups, rd = rows / groups_dot;
#define ITN_L(V) launch(layernorm_fwd_jvp_kernel<V>, grid, 256, 0, s, x, x_dot, mean, rstd, gamma, gamma_dot, beta_dot, y_dot, rows, rg, g_stride, rd, d_stride)
  switch (cols) {
    case 128: ITN_L(1); break;
    case 256: ITN_L(2); break;
    case 512: ITN_L(4); break;
    default: return set_error(ITN_ERR_UNSUPPORTED, "layernorm_fwd_jvp: cols must be 128/256/512, got %d", cols);
  }
#undef ITN_L
  return check_launch("layernorm_fwd_jvp_kernel");
}

extern "C" int itn_layernorm_bwd_jvp(const float* dy, const float* dy_dot, const float* x, const float* x_dot,
                                     const float* mean, const float* rstd, const float* gamma,
                                     const float* gamma_dot, float* dx_dot, float* gterm, long long rows, int cols,
                                     int groups, long long g_stride, int groups_dot, long long d_stride,
                                     void* stream) {
  ITN_REQUIRE(dy && x && mean && rstd && gamma && dx_dot, "layernorm_bwd_jvp: null pointer");
  ITN_REQUIRE(rows > 0 && groups > 0 && groups_dot > 0 && rows % groups == 0 && rows % groups_dot == 0,
              "layernorm_bwd_jvp: rows (%lld) must be a multiple of groups (%d, %d)", rows, groups, groups_dot);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const unsigned grid = (unsigned)((rows + 7) / 8);
  const long long rg = rows / groups, rd = rows / groups_dot;
#define ITN_L(V) launch(layernorm_bwd_jvp_kernel<V>, grid, 256, 0, s, dy, dy_dot, x, x_dot, mean, rstd, gamma, gamma_dot, dx_dot, gterm, rows, rg, g_stride, rd, d_stride)
  switch (cols) {
    case 128: ITN_L(1); break;
    case 256: ITN_L(2); break;
    case 512: ITN_L(4); break;
    default: return set_error(ITN_ERR_UNSUPPORTED, "layernorm_bwd_jvp: cols must be 128/256/512, got %d", cols);
  }
#undef ITN_L
  return check_launch("layernorm_bwd_jvp_kernel");
}

extern "C" int itn_softmax_bwd_jvp(const float* p, const float* p_dot, float* dp, float* dp_dot, long long rows,
                                   int cols, long long ld, float scale, void* stream) {
  ITN_REQUIRE(p && dp && dp_dot && rows > 0 && cols > 0 && ld >= cols, "softmax_bwd_jvp: bad arguments");
  ITN_REQUIRE((ld % 4) == 0 && ld >= ((cols + 3) / 4) * 4 && cols <= 128 * 20 && al16(p) && al16(dp) && al16(dp_dot) &&
                  (!p_dot || al16(p_dot)),
              "softmax_bwd_jvp: rows must be 16-byte aligned and own their padding (ld %% 4 == 0), cols <= 2560");
  const unsigned grid = (unsigned)((rows + 7) / 8);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define ITN_S(NV) launch(softmax_bwd_jvp_kernel<NV>, grid, 256, 0, st, p, p_dot, dp, dp_dot, rows, cols, ld, scale)
  if (cols <= 128) ITN_S(1);
  else if (cols <= 256) ITN_S(2);
  else if (cols <= 384) ITN_S(3);
  else if (cols <= 512) ITN_S(4);
  else if (cols <= 1024) ITN_S(8);
  else if (cols <= 2048) ITN_S(16);
  else ITN_S(20);
#undef ITN_S
  return check_launch("softmax_bwd_jvp_kernel");
}

extern "C" int itn_mask_mul(float* y, const float* ref, long long n, void* stream) {
  ITN_REQUIRE(y && ref && n > 0 && n % 4 == 0 && al16(y) && al16(ref), "mask_mul: need n %% 4 == 0 and 16-byte alignment");
  launch(mask_mul_kernel, ew_grid(n / 4), 256, 0, static_cast<cudaStream_t>(stream), y, ref, n / 4);
  return check_launch("mask_mul_kernel");
}

extern "C" int itn_mul_mask_u8(const float* x, const unsigned char* mask, float scale, float* out, long long n,
                               void* stream) {
  ITN_REQUIRE(x && mask && out && n > 0 && n % 4 == 0 && al16(x) && al16(out) && (((uintptr_t)mask) & 3) == 0,
              "mul_mask_u8: need n %% 4 == 0 and aligned pointers");
  launch(mul_mask_u8_kernel, ew_grid(n / 4), 256, 0, static_cast<cudaStream_t>(stream), x, mask, scale, out, n / 4);
  return check_launch("mul_mask_u8_kernel");
}

extern "C" int itn_gelu_grad_dual(const float* raw, const float* raw_dot, const float* aux, const float* aux_dot,
                                  float* y, float* y_dot, long long n, void* stream) {
  ITN_REQUIRE(raw && aux && y && n > 0 && n % 4 == 0, "gelu_grad_dual: need n %% 4 == 0");
  ITN_REQUIRE(al16(raw) && al16(aux) && al16(y) && al16(raw_dot) && al16(aux_dot) && al16(y_dot),
              "gelu_grad_dual: pointers must be 16-byte aligned");
  launch(gelu_grad_dual_kernel, ew_grid(n / 4), 256, 0, static_cast<cudaStream_t>(stream), raw, raw_dot, aux, aux_dot,
         y, y_dot, n / 4);
  return check_launch("gelu_grad_dual_kernel");
}

extern "C" int itn_sigmoid_bwd_jvp(const float* dy, const float* dy_dot, const float* y, const float* y_dot,
                                   float* out, long long n, void* stream) {
  ITN_REQUIRE(dy && y && out && n > 0, "sigmoid_bwd_jvp: bad arguments");
  launch(sigmoid_bwd_jvp_kernel, ew_grid(n), 256, 0, static_cast<cudaStream_t>(stream), dy, dy_dot, y, y_dot, out, n);
  return check_launch("sigmoid_bwd_jvp_kernel");
}

extern "C" int itn_l2norm_jvp(const float* x_dot, const float* nrm, const float* d, float* n_dot, float* d_dot,
                              int groups, int n, void* stream) {
  ITN_REQUIRE(x_dot && nrm && d && n_dot && d_dot && groups > 0 && n > 0, "l2norm_jvp: bad arguments");
  launch(l2norm_jvp_kernel, groups, 256, 0, static_cast<cudaStream_t>(stream), x_dot, nrm, d, n_dot, d_dot, n);
  return check_launch("l2norm_jvp_kernel");
}
