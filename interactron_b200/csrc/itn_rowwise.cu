// Row-wise HBM-bound kernels: LayerNorm fwd/bwd, softmax fwd/bwd, column sums.
// One warp per row, warp-shuffle reductions, coalesced (float4 where the row
// stride allows) accesses.  Contracts in include/interactron_b200.h.
#include "itn_common.cuh"

namespace itn {

// ------------------------------------------------------------- LayerNorm fwd
// cols = VPL * 128: lane owns float4 chunks lane, lane+32, ...
template <int VPL>
__global__ void __launch_bounds__(256)
layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                     const float* __restrict__ beta, float* __restrict__ y,
                     float* __restrict__ y_r, float* __restrict__ mean_out, float* __restrict__ rstd_out, long long rows,
                     long long rows_per_group, long long gb_stride, float eps,
                     const float* __restrict__ plus, float* __restrict__ y_plus, long long plus_elems,
                     long long plus_group, long long plus_gs) {
  pdl_wait();
  pdl_trigger();
  constexpr int cols = VPL * 128;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float4* xr = reinterpret_cast<const float4*>(x + row * cols);
  float4 v[VPL];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    v[i] = xr[lane + 32 * i];
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) * (1.0f / cols);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / cols) + eps);
  const long long g = row / rows_per_group;
  const float4* gr = reinterpret_cast<const float4*>(gamma + g * gb_stride);
  const float4* br = reinterpret_cast<const float4*>(beta + g * gb_stride);
  float4* yr = reinterpret_cast<float4*>(y + row * cols);
  float4* yrr = y_r ? reinterpret_cast<float4*>(y_r + row * cols) : nullptr;
  // second output y + plus (the `x + pos` / `tgt + query_pos` that feeds the next attention's q/k projection):
  // plus is a block of plus_elems values repeated inside each group of plus_group elements, one block per group
  const float4* pr = nullptr;
  float4* ypr = nullptr;
  if (y_plus) {
    const long long i0 = row * cols;
    pr = reinterpret_cast<const float4*>(plus + (i0 / plus_group) * plus_gs + i0 % plus_elems);
    ypr = reinterpret_cast<float4*>(y_plus + i0);
  }
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const float4 ga = gr[lane + 32 * i], be = br[lane + 32 * i];
    float4 o;
    o.x = (v[i].x - mean) * rstd * ga.x + be.x;
    o.y = (v[i].y - mean) * rstd * ga.y + be.y;
    o.z = (v[i].z - mean) * rstd * ga.z + be.z;
    o.w = (v[i].w - mean) * rstd * ga.w + be.w;
    yr[lane + 32 * i] = o;
    if (yrr) yrr[lane + 32 * i] = make_float4(rn_tf32(o.x), rn_tf32(o.y), rn_tf32(o.z), rn_tf32(o.w));
    if (ypr) {
      const float4 pv = pr[lane + 32 * i];
      ypr[lane + 32 * i] = make_float4(o.x + pv.x, o.y + pv.y, o.z + pv.z, o.w + pv.w);
    }
  }
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
}

// ------------------------------------------------------------- LayerNorm bwd
template <int VPL>
__global__ void __launch_bounds__(256)
layernorm_bwd_dx_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                        const float* __restrict__ mean, const float* __restrict__ rstd,
                        const float* __restrict__ gamma, float* __restrict__ dx,
                        float* __restrict__ dx_r, long long rows, long long rows_per_group,
                        long long gb_stride) {
  pdl_wait();
  pdl_trigger();
  constexpr int cols = VPL * 128;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float mu = mean[row], rs = rstd[row];
  const long long g = row / rows_per_group;
  const float4* xr = reinterpret_cast<const float4*>(x + row * cols);
  const float4* dr = reinterpret_cast<const float4*>(dy + row * cols);
  const float4* gr = reinterpret_cast<const float4*>(gamma + g * gb_stride);
  float4 xh[VPL], dg[VPL];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const float4 xv = xr[lane + 32 * i], dv = dr[lane + 32 * i], ga = gr[lane + 32 * i];
    xh[i].x = (xv.x - mu) * rs; xh[i].y = (xv.y - mu) * rs;
    xh[i].z = (xv.z - mu) * rs; xh[i].w = (xv.w - mu) * rs;
    dg[i].x = dv.x * ga.x; dg[i].y = dv.y * ga.y; dg[i].z = dv.z * ga.z; dg[i].w = dv.w * ga.w;
    s1 += (dg[i].x + dg[i].y) + (dg[i].z + dg[i].w);
    s2 += (dg[i].x * xh[i].x + dg[i].y * xh[i].y) + (dg[i].z * xh[i].z + dg[i].w * xh[i].w);
  }
  const float m1 = warp_sum(s1) * (1.0f / cols);
  const float m2 = warp_sum(s2) * (1.0f / cols);
  float4* or_ = reinterpret_cast<float4*>(dx + row * cols);
  float4* orr = dx_r ? reinterpret_cast<float4*>(dx_r + row * cols) : nullptr;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    float4 o;
    o.x = rs * (dg[i].x - m1 - xh[i].x * m2);
    o.y = rs * (dg[i].y - m1 - xh[i].y * m2);
    o.z = rs * (dg[i].z - m1 - xh[i].z * m2);
    o.w = rs * (dg[i].w - m1 - xh[i].w * m2);
    or_[lane + 32 * i] = o;
    if (orr) orr[lane + 32 * i] = make_float4(rn_tf32(o.x), rn_tf32(o.y), rn_tf32(o.z), rn_tf32(o.w));
  }
}

// dgamma[g,c] = sum_r dy*xhat, dbeta[g,c] = sum_r dy over the rows of group g.
// Block = 32 columns x 32 row-lanes; fixed summation order (deterministic).
__global__ void __launch_bounds__(1024)
layernorm_bwd_gb_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                        const float* __restrict__ mean, const float* __restrict__ rstd,
                        float* __restrict__ dgamma, float* __restrict__ dbeta,
                        long long rows_per_group, int cols, long long dgb_stride) {
  pdl_wait();
  pdl_trigger();
  __shared__ float sg[32][33];
  __shared__ float sb[32][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int g = blockIdx.y;
  const long long r0 = (long long)g * rows_per_group;
  float ag = 0.f, ab = 0.f;
  if (c < cols) {
    for (long long r = threadIdx.y; r < rows_per_group; r += 32) {
      const long long row = r0 + r;
      const float d = dy[row * cols + c];
      ag += d * (x[row * cols + c] - mean[row]) * rstd[row];
      ab += d;
    }
  }
  sg[threadIdx.y][threadIdx.x] = ag;
  sb[threadIdx.y][threadIdx.x] = ab;
  __syncthreads();
  if (threadIdx.y == 0 && c < cols) {
    float tg = 0.f, tb = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      tg += sg[i][threadIdx.x];
      tb += sb[i][threadIdx.x];
    }
    if (dgamma) dgamma[(long long)g * dgb_stride + c] = tg;
    if (dbeta) dbeta[(long long)g * dgb_stride + c] = tb;
  }
}


// One launch for the whole LayerNorm backward: dx, the per-group column sums dgamma = sum dy*xhat and
// dbeta = sum dy, and (optionally) dxsum = sum dx - the bias gradient of the linear layer that feeds the
// residual this LayerNorm normalises (post-norm blocks: d(x + sublayer(x)) = dx goes to both), which
// otherwise costs a colsum launch re-reading dx.  CTA (chunk, g) walks `chunk_rows` rows of group g, one
// warp per row, column partials in registers; the CTA's partial goes to `partials`, and the CTA that
// arrives last at the group's counter adds the partials up in chunk order: fixed summation order
// (deterministic), no second launch.  The counter is reset by that CTA (workspace is reusable as is).
template <int VPL>
__global__ void __launch_bounds__(256)
layernorm_bwd_fused_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                           const float* __restrict__ mean, const float* __restrict__ rstd,
                           const float* __restrict__ gamma, float* __restrict__ dx,
                           float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dxsum,
                           float* __restrict__ partials, unsigned* __restrict__ counters,
                           long long rows_per_group, int chunk_rows, long long gb_stride,
                           long long dgb_stride, long long dxsum_stride) {
  pdl_wait();
  pdl_trigger();
  constexpr int cols = VPL * 128;
  __shared__ float red[3 * cols];
  __shared__ bool last;
  const int chunk = blockIdx.x, chunks = gridDim.x, g = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long r_begin = (long long)chunk * chunk_rows;
  const long long r_end = r_begin + chunk_rows < rows_per_group ? r_begin + chunk_rows : rows_per_group;
  const float4* gr = reinterpret_cast<const float4*>(gamma + g * gb_stride);
  float4 ga[VPL], ag[VPL], ab[VPL], ax[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    ga[i] = gr[lane + 32 * i];
    ag[i] = ab[i] = ax[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (long long r = r_begin + warp; r < r_end; r += 8) {
    const long long row = (long long)g * rows_per_group + r;
    const float mu = mean[row], rs = rstd[row];
    const float4* xr = reinterpret_cast<const float4*>(x + row * cols);
    const float4* dr = reinterpret_cast<const float4*>(dy + row * cols);
    float4 xh[VPL], dg[VPL];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const float4 xv = xr[lane + 32 * i], dv = dr[lane + 32 * i];
      xh[i].x = (xv.x - mu) * rs; xh[i].y = (xv.y - mu) * rs;
      xh[i].z = (xv.z - mu) * rs; xh[i].w = (xv.w - mu) * rs;
      dg[i].x = dv.x * ga[i].x; dg[i].y = dv.y * ga[i].y; dg[i].z = dv.z * ga[i].z; dg[i].w = dv.w * ga[i].w;
      s1 += (dg[i].x + dg[i].y) + (dg[i].z + dg[i].w);
      s2 += (dg[i].x * xh[i].x + dg[i].y * xh[i].y) + (dg[i].z * xh[i].z + dg[i].w * xh[i].w);
      ag[i].x += dv.x * xh[i].x; ag[i].y += dv.y * xh[i].y; ag[i].z += dv.z * xh[i].z; ag[i].w += dv.w * xh[i].w;
      ab[i].x += dv.x; ab[i].y += dv.y; ab[i].z += dv.z; ab[i].w += dv.w;
    }
    const float m1 = warp_sum(s1) * (1.0f / cols);
    const float m2 = warp_sum(s2) * (1.0f / cols);
    float4* or_ = reinterpret_cast<float4*>(dx + row * cols);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      float4 o;
      o.x = rs * (dg[i].x - m1 - xh[i].x * m2);
      o.y = rs * (dg[i].y - m1 - xh[i].y * m2);
      o.z = rs * (dg[i].z - m1 - xh[i].z * m2);
      o.w = rs * (dg[i].w - m1 - xh[i].w * m2);
      or_[lane + 32 * i] = o;
      ax[i].x += o.x; ax[i].y += o.y; ax[i].z += o.z; ax[i].w += o.w;
    }
  }
  // warps add their column partials into shared memory one after the other (fixed order)
  float4* red4 = reinterpret_cast<float4*>(red);
  for (int w = 0; w < 8; ++w) {
    if (warp == w) {
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const int c4 = lane + 32 * i;
        if (w == 0) {
          red4[c4] = ag[i];
          red4[cols / 4 + c4] = ab[i];
          red4[2 * (cols / 4) + c4] = ax[i];
        } else {
          float4 t = red4[c4];
          red4[c4] = make_float4(t.x + ag[i].x, t.y + ag[i].y, t.z + ag[i].z, t.w + ag[i].w);
          t = red4[cols / 4 + c4];
          red4[cols / 4 + c4] = make_float4(t.x + ab[i].x, t.y + ab[i].y, t.z + ab[i].z, t.w + ab[i].w);
          t = red4[2 * (cols / 4) + c4];
          red4[2 * (cols / 4) + c4] = make_float4(t.x + ax[i].x, t.y + ax[i].y, t.z + ax[i].z, t.w + ax[i].w);
        }
      }
    }
    __syncthreads();
  }
  float* outs[3] = {dgamma ? dgamma + (long long)g * dgb_stride : nullptr, dbeta ? dbeta + (long long)g * dgb_stride : nullptr,
                    dxsum ? dxsum + (long long)g * dxsum_stride : nullptr};
  if (chunks == 1) {            // nothing to combine
    for (int idx = threadIdx.x; idx < 3 * cols; idx += 256) {
      float* o = outs[idx / cols];
      if (o) o[idx % cols] = red[idx];
    }
    return;
  }
  float* mine = partials + ((long long)g * chunks + chunk) * (3 * cols);
  for (int idx = threadIdx.x; idx < 3 * cols; idx += 256) mine[idx] = red[idx];
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned old = atomicAdd(&counters[g], 1u);
    last = (old == (unsigned)chunks - 1u);
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  const float* gp = partials + (long long)g * chunks * (3 * cols);
  for (int idx = threadIdx.x; idx < 3 * cols; idx += 256) {
    float* o = outs[idx / cols];
    if (!o) continue;
    float t = 0.f;
    for (int ch = 0; ch < chunks; ++ch) t += __ldcg(gp + (long long)ch * (3 * cols) + idx);
    o[idx % cols] = t;
  }
  if (threadIdx.x == 0) counters[g] = 0u;
}

// ------------------------------------------------------------------ softmax
// One warp per row, online max/sum pass then a normalising pass (the row is
// re-read from L1/L2, never from HBM: rows are <= 8 KB).
__global__ void __launch_bounds__(256)
softmax_fwd_kernel(float* __restrict__ s, long long rows, int cols, long long ld, float scale,
                   const unsigned char* __restrict__ key_mask, long long rows_per_mask,
                   int round_out) {
  pdl_wait();
  pdl_trigger();
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  float* r = s + row * ld;
  const unsigned char* mk = key_mask ? key_mask + (row / rows_per_mask) * cols : nullptr;
  float m = -INFINITY, l = 0.f;
  for (int c = lane; c < cols; c += 32) {
    float v = r[c] * scale;
    if (mk && mk[c]) v = -INFINITY;
    if (v > m) {
      l = l * __expf(m - v) + 1.0f;
      m = v;
    } else if (v != -INFINITY) {
      l += __expf(v - m);
    }
  }
  const float gm = warp_max(m);
  l = (m == -INFINITY) ? 0.f : l * __expf(m - gm);
  const float inv = 1.0f / warp_sum(l);
  for (int c = lane; c < cols; c += 32) {
    float v = r[c] * scale;
    if (mk && mk[c]) v = -INFINITY;
    const float o = __expf(v - gm) * inv;
    r[c] = round_out ? rn_tf32(o) : o;
  }
}

__global__ void __launch_bounds__(256)
softmax_bwd_kernel(const float* __restrict__ p, float* __restrict__ dp, long long rows, int cols,
                   long long ld, float scale, int round_out) {
  pdl_wait();
  pdl_trigger();
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* pr = p + row * ld;
  float* dr = dp + row * ld;
  float acc = 0.f;
  for (int c = lane; c < cols; c += 32) acc += pr[c] * dr[c];
  const float dot = warp_sum(acc);
  for (int c = lane; c < cols; c += 32) {
    const float o = scale * pr[c] * (dr[c] - dot);
    dr[c] = round_out ? rn_tf32(o) : o;
  }
}

// Register-resident variants: a warp keeps its whole row (<= 128*NV columns) in registers as NV
// float4 per lane, so every element is read from memory once with all loads in flight at
// once, and written once (the generic kernels above re-read the row and carry a serial exp chain
// in the online pass: 2.4 TB/s on the 361-column encoder rows, profiles/README.md).
// Requires 16-byte aligned rows (ld % 4 == 0) that own their padding up to a multiple of 4 columns;
// the pad columns are written as zeros.
template <int NV>
__global__ void __launch_bounds__(256)
softmax_fwd_reg_kernel(float* __restrict__ s, long long rows, int cols, long long ld, float scale,
                       const unsigned char* __restrict__ key_mask, long long rows_per_mask,
                       int round_out) {
  pdl_wait();
  pdl_trigger();
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  float4* r4 = reinterpret_cast<float4*>(s + row * ld);
  const unsigned char* mk = key_mask ? key_mask + (row / rows_per_mask) * cols : nullptr;
  const int nvec = (cols + 3) >> 2;
  float v[NV][4];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int j = lane + 32 * i;
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    if (j < nvec) t = r4[j];
    v[i][0] = t.x; v[i][1] = t.y; v[i][2] = t.z; v[i][3] = t.w;
  }
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c0 = 4 * (lane + 32 * i);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = c0 + k;
      const bool live = c < cols && !(mk && mk[c]);
      v[i][k] = live ? v[i][k] * scale : -INFINITY;
      m = fmaxf(m, v[i][k]);
    }
  }
  m = warp_max(m);
  float l = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      v[i][k] = __expf(v[i][k] - m);      // exp(-inf) = 0 for masked / pad columns
      l += v[i][k];
    }
  }
  const float inv = 1.0f / warp_sum(l);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int j = lane + 32 * i;
    if (j < nvec) {
      float4 o = make_float4(v[i][0] * inv, v[i][1] * inv, v[i][2] * inv, v[i][3] * inv);
      if (round_out) o = make_float4(rn_tf32(o.x), rn_tf32(o.y), rn_tf32(o.z), rn_tf32(o.w));
      r4[j] = o;
    }
  }
}

template <int NV>
__global__ void __launch_bounds__(256)
softmax_bwd_reg_kernel(const float* __restrict__ p, float* __restrict__ dp, long long rows, int cols,
                       long long ld, float scale, int round_out) {
  pdl_wait();
  pdl_trigger();
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float4* p4 = reinterpret_cast<const float4*>(p + row * ld);
  float4* d4 = reinterpret_cast<float4*>(dp + row * ld);
  const int nvec = (cols + 3) >> 2;
  float pv[NV][4], dv[NV][4];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int j = lane + 32 * i;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (j < nvec) {
      a = p4[j];
      b = d4[j];
    }
    pv[i][0] = a.x; pv[i][1] = a.y; pv[i][2] = a.z; pv[i][3] = a.w;
    dv[i][0] = b.x; dv[i][1] = b.y; dv[i][2] = b.z; dv[i][3] = b.w;
  }
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c0 = 4 * (lane + 32 * i);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (c0 + k >= cols) pv[i][k] = 0.f, dv[i][k] = 0.f;    // pad columns of the last chunk
      acc = fmaf(pv[i][k], dv[i][k], acc);
    }
  }
  const float dot = warp_sum(acc);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int j = lane + 32 * i;
    if (j < nvec) {
      float4 o = make_float4(scale * pv[i][0] * (dv[i][0] - dot), scale * pv[i][1] * (dv[i][1] - dot),
                             scale * pv[i][2] * (dv[i][2] - dot), scale * pv[i][3] * (dv[i][3] - dot));
      if (round_out) o = make_float4(rn_tf32(o.x), rn_tf32(o.y), rn_tf32(o.z), rn_tf32(o.w));
      d4[j] = o;
    }
  }
}

// rows the register-resident kernels can take: 16-byte aligned, padding owned by the row
static inline bool softmax_rows_vectorisable(const void* a, const void* b, int cols, long long ld) {
  return (ld % 4) == 0 && ld >= ((cols + 3) / 4) * 4 && cols <= 128 * 20 &&
         (((uintptr_t)a | (uintptr_t)b) & 15) == 0;
}

// ------------------------------------------------------------------- colsum
__global__ void __launch_bounds__(1024)
colsum_kernel(const float* __restrict__ x, float* __restrict__ out, long long rows, int cols,
              long long ld, long long out_stride) {
  pdl_wait();
  pdl_trigger();
  __shared__ float sm[32][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int g = blockIdx.y;
  const float* xg = x + (long long)g * rows * ld;
  float a = 0.f;
  if (c < cols)
    for (long long r = threadIdx.y; r < rows; r += 32) a += xg[r * ld + c];
  sm[threadIdx.y][threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.y == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) t += sm[i][threadIdx.x];
    out[(long long)g * out_stride + c] = t;
  }
}

}  // namespace itn

using namespace itn;

static int layernorm_fwd_launch(const float* x, const float* gamma, const float* beta, float* y,
                                float* y_r, float* mean, float* rstd, long long rows, int cols, int groups,
                                long long gb_stride, float eps, const float* plus, float* y_plus,
                                long long plus_elems, long long plus_group, long long plus_gs, void* stream) {
  ITN_REQUIRE(x && gamma && beta && y, "layernorm_fwd: null pointer");
  ITN_REQUIRE(rows > 0 && groups > 0 && rows % groups == 0,
              "layernorm_fwd: rows (%lld) must be a positive multiple of groups (%d)", rows, groups);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const unsigned grid = (unsigned)((rows + 7) / 8);
  const long long rpg = rows / groups;
#define ITN_LNF(V) launch(layernorm_fwd_kernel<V>, grid, 256, 0, s, x, gamma, beta, y, y_r, mean, rstd, rows, rpg, gb_stride, eps, \
                          plus, y_plus, plus_elems, plus_group, plus_gs)
  switch (cols) {
    case 128: ITN_LNF(1); break;
    case 256: ITN_LNF(2); break;
    case 512: ITN_LNF(4); break;
    case 1024: ITN_LNF(8); break;
    default: return set_error(ITN_ERR_UNSUPPORTED, "layernorm_fwd: cols must be 128/256/512/1024, got %d", cols);
  }
#undef ITN_LNF
  return check_launch("layernorm_fwd_kernel");
}

extern "C" int itn_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* y,
                                 float* y_r, float* mean, float* rstd, long long rows, int cols, int groups,
                                 long long gb_stride, float eps, void* stream) {
  return layernorm_fwd_launch(x, gamma, beta, y, y_r, mean, rstd, rows, cols, groups, gb_stride, eps, nullptr, nullptr,
                              1, 1, 0, stream);
}

extern "C" int itn_layernorm_fwd_plus(const float* x, const float* gamma, const float* beta, float* y,
                                      float* mean, float* rstd, long long rows, int cols, int groups,
                                      long long gb_stride, float eps, const float* plus, float* y_plus,
                                      long long plus_elems, long long plus_group, long long plus_group_stride,
                                      void* stream) {
  ITN_REQUIRE(plus && y_plus, "layernorm_fwd_plus: null pointer");
  ITN_REQUIRE(plus_elems > 0 && plus_elems % cols == 0 && plus_group > 0 && plus_group % plus_elems == 0 &&
              (rows * (long long)cols) % plus_group == 0 && plus_group_stride % 4 == 0 &&
              (reinterpret_cast<uintptr_t>(plus) & 15) == 0 && (reinterpret_cast<uintptr_t>(y_plus) & 15) == 0,
              "layernorm_fwd_plus: the added block must be whole rows (plus_elems %% cols == 0) tiling each group, 16-byte aligned");
  return layernorm_fwd_launch(x, gamma, beta, y, nullptr, mean, rstd, rows, cols, groups, gb_stride, eps, plus, y_plus,
                              plus_elems, plus_group, plus_group_stride, stream);
}

extern "C" int itn_layernorm_bwd(const float* dy, const float* x, const float* mean,
                                 const float* rstd, const float* gamma, float* dx, float* dx_r,
                                 float* dgamma, float* dbeta, long long rows, int cols, int groups,
                                 long long gb_stride, long long dgb_stride, void* stream) {
  ITN_REQUIRE(dy && x && mean && rstd && gamma && dx, "layernorm_bwd: null pointer");
  ITN_REQUIRE(rows > 0 && groups > 0 && rows % groups == 0,
              "layernorm_bwd: rows (%lld) must be a positive multiple of groups (%d)", rows, groups);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const unsigned grid = (unsigned)((rows + 7) / 8);
  const long long rpg = rows / groups;
  switch (cols) {
    case 128: launch(layernorm_bwd_dx_kernel<1>, grid, 256, 0, s, dy, x, mean, rstd, gamma, dx, dx_r, rows, rpg, gb_stride); break;
    case 256: launch(layernorm_bwd_dx_kernel<2>, grid, 256, 0, s, dy, x, mean, rstd, gamma, dx, dx_r, rows, rpg, gb_stride); break;
    case 512: launch(layernorm_bwd_dx_kernel<4>, grid, 256, 0, s, dy, x, mean, rstd, gamma, dx, dx_r, rows, rpg, gb_stride); break;
    case 1024: launch(layernorm_bwd_dx_kernel<8>, grid, 256, 0, s, dy, x, mean, rstd, gamma, dx, dx_r, rows, rpg, gb_stride); break;
    default: return set_error(ITN_ERR_UNSUPPORTED, "layernorm_bwd: cols must be 128/256/512/1024, got %d", cols);
  }
  int rc = check_launch("layernorm_bwd_dx_kernel");
  if (rc) return rc;
  if (dgamma || dbeta) {
    dim3 g2((cols + 31) / 32, groups);
    launch(layernorm_bwd_gb_kernel, g2, dim3(32, 32), 0, s, dy, x, mean, rstd, dgamma, dbeta, rpg, cols, dgb_stride);
    rc = check_launch("layernorm_bwd_gb_kernel");
  }
  return rc;
}


// chunks of rows per group so that the grid has ~2 CTAs per SM and every warp still gets >= 1 row
static int ln_bwd_chunks(long long rpg, int groups) {
  long long want = (296 + groups - 1) / groups;
  const long long cap = (rpg + 7) / 8;
  if (want > cap) want = cap;
  if (want > 512) want = 512;
  return want < 1 ? 1 : (int)want;
}

extern "C" long long itn_layernorm_bwd_fused_workspace(long long rows, int cols, int groups) {
  if (rows <= 0 || groups <= 0 || rows % groups) return 0;
  const int chunks = ln_bwd_chunks(rows / groups, groups);
  return 4LL * groups * chunks * 3 * cols + 4LL * ((groups + 63) / 64 * 64);
}

extern "C" int itn_layernorm_bwd_fused(const float* dy, const float* x, const float* mean, const float* rstd,
                                       const float* gamma, float* dx, float* dgamma, float* dbeta, float* dxsum,
                                       long long rows, int cols, int groups, long long gb_stride,
                                       long long dgb_stride, long long dxsum_stride, void* workspace,
                                       long long workspace_bytes, void* stream) {
  ITN_REQUIRE(dy && x && mean && rstd && gamma && dx, "layernorm_bwd_fused: null pointer");
  ITN_REQUIRE(rows > 0 && groups > 0 && rows % groups == 0,
              "layernorm_bwd_fused: rows (%lld) must be a positive multiple of groups (%d)", rows, groups);
  ITN_REQUIRE(groups <= 65535, "layernorm_bwd_fused: too many groups (%d)", groups);
  const long long need = itn_layernorm_bwd_fused_workspace(rows, cols, groups);
  ITN_REQUIRE(workspace && workspace_bytes >= need && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0,
              "layernorm_bwd_fused: workspace of %lld bytes (16-byte aligned, counters zeroed once) needed, got %lld",
              need, workspace_bytes);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long rpg = rows / groups;
  const int chunks = ln_bwd_chunks(rpg, groups);
  const int chunk_rows = (int)((rpg + chunks - 1) / chunks);
  // counters first (zeroed by the caller once; the kernel leaves them zero), partial sums behind them
  unsigned* counters = static_cast<unsigned*>(workspace);
  float* partials = reinterpret_cast<float*>(counters + (groups + 63) / 64 * 64);
  dim3 grid(chunks, groups);
#define ITN_LNB(V) launch(layernorm_bwd_fused_kernel<V>, grid, 256, 0, s, dy, x, mean, rstd, gamma, dx, dgamma, dbeta, dxsum, \
                          partials, counters, rpg, chunk_rows, gb_stride, dgb_stride, dxsum_stride)
  switch (cols) {
    case 128: ITN_LNB(1); break;
    case 256: ITN_LNB(2); break;
    case 512: ITN_LNB(4); break;
    default: return set_error(ITN_ERR_UNSUPPORTED, "layernorm_bwd_fused: cols must be 128/256/512, got %d", cols);
  }
#undef ITN_LNB
  return check_launch("layernorm_bwd_fused_kernel");
}

extern "C" int itn_softmax_fwd(float* sc, long long rows, int cols, long long ld, float scale,
                               const unsigned char* key_mask, long long rows_per_mask,
                               int round_out, void* stream) {
  ITN_REQUIRE(sc && rows > 0 && cols > 0 && ld >= cols, "softmax_fwd: bad arguments");
  ITN_REQUIRE(!key_mask || rows_per_mask > 0, "softmax_fwd: rows_per_mask must be > 0 with a mask");
  const unsigned grid = (unsigned)((rows + 7) / 8);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (softmax_rows_vectorisable(sc, nullptr, cols, ld)) {
#define ITN_SOFTMAX_FWD(NV) launch(softmax_fwd_reg_kernel<NV>, grid, 256, 0, st, sc, rows, cols, ld, scale, key_mask, rows_per_mask, round_out)
    if (cols <= 128) ITN_SOFTMAX_FWD(1);
    else if (cols <= 256) ITN_SOFTMAX_FWD(2);
    else if (cols <= 384) ITN_SOFTMAX_FWD(3);
    else if (cols <= 512) ITN_SOFTMAX_FWD(4);
    else if (cols <= 1024) ITN_SOFTMAX_FWD(8);
    else if (cols <= 2048) ITN_SOFTMAX_FWD(16);
    else ITN_SOFTMAX_FWD(20);
#undef ITN_SOFTMAX_FWD
    return check_launch("softmax_fwd_reg_kernel");
  }
  launch(softmax_fwd_kernel, grid, 256, 0, st, sc, rows, cols, ld, scale, key_mask, rows_per_mask, round_out);
  return check_launch("softmax_fwd_kernel");
}

extern "C" int itn_softmax_bwd(const float* p, float* dp, long long rows, int cols, long long ld,
                               float scale, int round_out, void* stream) {
  ITN_REQUIRE(p && dp && rows > 0 && cols > 0 && ld >= cols, "softmax_bwd: bad arguments");
  const unsigned grid = (unsigned)((rows + 7) / 8);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (softmax_rows_vectorisable(p, dp, cols, ld)) {
#define ITN_SOFTMAX_BWD(NV) launch(softmax_bwd_reg_kernel<NV>, grid, 256, 0, st, p, dp, rows, cols, ld, scale, round_out)
    if (cols <= 128) ITN_SOFTMAX_BWD(1);
    else if (cols <= 256) ITN_SOFTMAX_BWD(2);
    else if (cols <= 384) ITN_SOFTMAX_BWD(3);
    else if (cols <= 512) ITN_SOFTMAX_BWD(4);
    else if (cols <= 1024) ITN_SOFTMAX_BWD(8);
    else if (cols <= 2048) ITN_SOFTMAX_BWD(16);
    else ITN_SOFTMAX_BWD(20);
#undef ITN_SOFTMAX_BWD
    return check_launch("softmax_bwd_reg_kernel");
  }
  launch(softmax_bwd_kernel, grid, 256, 0, st, p, dp, rows, cols, ld, scale, round_out);
  return check_launch("softmax_bwd_kernel");
}

extern "C" int itn_colsum(const float* x, float* out, int groups, long long rows, int cols,
                          long long ld, long long out_stride, void* stream) {
  ITN_REQUIRE(x && out && groups > 0 && rows > 0 && cols > 0 && ld >= cols, "colsum: bad arguments");
  dim3 grid((cols + 31) / 32, groups);
  launch(colsum_kernel, grid, dim3(32, 32), 0, static_cast<cudaStream_t>(stream), x, out, rows, cols, ld, out_stride);
  return check_launch("colsum_kernel");
}
