// Element-wise HBM-bound kernels: fused fast-weight SGD step, add, strided copy,
// sigmoid fwd/bwd, learned-loss L2 norm, DETR sine position embedding.
// 128-bit accesses, grid-stride loops sized in multiples of the SM count.
#include "itn_common.cuh"
#include "itn_philox.cuh"

namespace itn {

constexpr int kSMs = 148;

static inline unsigned grid_for(long long work_items, int threads, int max_ctas_per_sm = 8) {
  long long blocks = (work_items + threads - 1) / threads;
  const long long cap = (long long)kSMs * max_ctas_per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

// theta_out = theta - clip(lr*g).  n4 = n/4 vector part, scalar tail handled by block 0.
__global__ void __launch_bounds__(256)
sgd_clip_update_kernel(const float* __restrict__ theta, long long theta_stride,
                       const float* __restrict__ g, float* __restrict__ out,
                       float* __restrict__ out_r, unsigned char* __restrict__ mask, long long n,
                       float lr, float clip) {
  pdl_wait();
  pdl_trigger();
  const int grp = blockIdx.y;
  const float* th = theta + grp * theta_stride;
  const float* gg = g + (long long)grp * n;
  float* oo = out + (long long)grp * n;
  float* orr = out_r ? out_r + (long long)grp * n : nullptr;
  unsigned char* mm = mask ? mask + (long long)grp * n : nullptr;
  const long long n4 = n >> 2;
  const float4* th4 = reinterpret_cast<const float4*>(th);
  const float4* g4 = reinterpret_cast<const float4*>(gg);
  float4* o4 = reinterpret_cast<float4*>(oo);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 t = th4[i];
    const float4 gv = __ldcs(g4 + i);
    float4 st, o;
    st.x = lr * gv.x; st.y = lr * gv.y; st.z = lr * gv.z; st.w = lr * gv.w;
    o.x = t.x - fminf(fmaxf(st.x, -clip), clip);
    o.y = t.y - fminf(fmaxf(st.y, -clip), clip);
    o.z = t.z - fminf(fmaxf(st.z, -clip), clip);
    o.w = t.w - fminf(fmaxf(st.w, -clip), clip);
    o4[i] = o;
    if (orr)
      reinterpret_cast<float4*>(orr)[i] =
          make_float4(rn_tf32(o.x), rn_tf32(o.y), rn_tf32(o.z), rn_tf32(o.w));
    if (mm) {
      uchar4 m;
      m.x = fabsf(st.x) <= clip; m.y = fabsf(st.y) <= clip;
      m.z = fabsf(st.z) <= clip; m.w = fabsf(st.w) <= clip;
      reinterpret_cast<uchar4*>(mm)[i] = m;
    }
  }
  if (blockIdx.x == 0) {
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) {
      const float st = lr * gg[i];
      const float o = th[i] - fminf(fmaxf(st, -clip), clip);
      oo[i] = o;
      if (orr) orr[i] = rn_tf32(o);
      if (mm) mm[i] = fabsf(st) <= clip;
    }
  }
}

__global__ void __launch_bounds__(256)
add_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
           long long n, long long b_elems, long long a_group, long long b_group_stride,
           int round_out) {
  pdl_wait();
  pdl_trigger();
  // all of n, b_elems, a_group, b_group_stride are multiples of 4 and pointers 16-byte aligned
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long n4 = n >> 2;
  const float4* a4 = reinterpret_cast<const float4*>(a);
  float4* o4 = reinterpret_cast<float4*>(out);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const long long e = i << 2;
    const long long bi = (e / a_group) * b_group_stride + e % b_elems;
    const float4 x = a4[i];
    const float4 y = *reinterpret_cast<const float4*>(b + bi);
    float4 o = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
    if (round_out) o = make_float4(rn_tf32(o.x), rn_tf32(o.y), rn_tf32(o.z), rn_tf32(o.w));
    o4[i] = o;
  }
}

__global__ void __launch_bounds__(256)
copy2d_kernel(const float* __restrict__ src, long long lds, float* __restrict__ dst, long long ldd,
              long long rows, int cols, int round_out) {
  pdl_wait();
  pdl_trigger();
  const long long total = rows * cols;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long r = i / cols;
    const int c = (int)(i - r * cols);
    const float v = src[r * lds + c];
    dst[r * ldd + c] = round_out ? rn_tf32(v) : v;
  }
}

// 32x32 tiles through padded shared memory: coalesced reads and writes.
__global__ void __launch_bounds__(256)
transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols,
                 long long sgs, long long dgs) {
  pdl_wait();
  pdl_trigger();
  __shared__ float t[32][33];
  const float* s = src + blockIdx.z * sgs;
  float* d = dst + blockIdx.z * dgs;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int r = r0 + ty + i, c = c0 + tx;
    if (r < rows && c < cols) t[ty + i][tx] = s[(long long)r * cols + c];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int c = c0 + ty + i, r = r0 + tx;
    if (r < rows && c < cols) d[(long long)c * rows + r] = t[tx][ty + i];
  }
}

__global__ void __launch_bounds__(256)
round_tf32_kernel(const float* __restrict__ src, float* __restrict__ dst, long long n) {
  pdl_wait();
  pdl_trigger();
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long n4 = n >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = reinterpret_cast<const float4*>(src)[i];
    reinterpret_cast<float4*>(dst)[i] =
        make_float4(rn_tf32(v.x), rn_tf32(v.y), rn_tf32(v.z), rn_tf32(v.w));
  }
  if (blockIdx.x == 0)
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) dst[i] = rn_tf32(src[i]);
}

__global__ void __launch_bounds__(256)
sigmoid_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long long n) {
  pdl_wait();
  pdl_trigger();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    y[i] = 1.0f / (1.0f + expf(-x[i]));
}

__global__ void __launch_bounds__(256)
sigmoid_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                   float* __restrict__ dx, long long n) {
  pdl_wait();
  pdl_trigger();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float s = y[i];
    dx[i] = dy[i] * s * (1.0f - s);
  }
}

// One block per group: loss = sqrt(sum x^2); dx = x / loss.
__global__ void __launch_bounds__(256)
l2norm_fwd_bwd_kernel(const float* __restrict__ x, float* __restrict__ loss,
                      float* __restrict__ dx, int n) {
  pdl_wait();
  pdl_trigger();
  __shared__ float sm[8];
  const float* xg = x + (long long)blockIdx.x * n;
  float a = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) a += xg[i] * xg[i];
  a = warp_sum(a);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = a;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) t += sm[i];
  const float nrm = sqrtf(t);
  if (threadIdx.x == 0 && loss) loss[blockIdx.x] = nrm;
  if (dx) {
    const float inv = 1.0f / nrm;
    float* dg = dx + (long long)blockIdx.x * n;
    for (int i = threadIdx.x; i < n; i += 256) dg[i] = xg[i] * inv;
  }
}

// One block per frame.  y_embed/x_embed = cumulative counts of unpadded cells,
// normalised by the last row/column (+1e-6) times 2*pi; channel c of an axis is
// sin (c even) / cos (c odd) of embed / 10000^(2*(c/2)/feats); y channels first.
__global__ void __launch_bounds__(256)
pos_embed_sine_kernel(const unsigned char* __restrict__ mask, float* __restrict__ pos, int h, int w,
                      int feats) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float sh[];
  float* ye = sh;            // [h*w]
  float* xe = sh + h * w;    // [h*w]
  const unsigned char* m = mask + (long long)blockIdx.x * h * w;
  for (int j = threadIdx.x; j < w; j += blockDim.x) {
    float c = 0.f;
    for (int i = 0; i < h; ++i) {
      c += m[i * w + j] ? 0.f : 1.f;
      ye[i * w + j] = c;
    }
  }
  for (int i = threadIdx.x; i < h; i += blockDim.x) {
    float c = 0.f;
    for (int j = 0; j < w; ++j) {
      c += m[i * w + j] ? 0.f : 1.f;
      xe[i * w + j] = c;
    }
  }
  __syncthreads();
  const float two_pi = 6.283185307179586f;
  const int D = 2 * feats;
  float* out = pos + (long long)blockIdx.x * h * w * D;
  for (int idx = threadIdx.x; idx < h * w * D; idx += blockDim.x) {
    const int t = idx / D, c = idx - t * D;
    const int i = t / w, j = t - i * w;
    const bool is_y = c < feats;
    const int cc = is_y ? c : c - feats;
    const float e = is_y ? ye[t] / (ye[(h - 1) * w + j] + 1e-6f) * two_pi
                         : xe[t] / (xe[i * w + (w - 1)] + 1e-6f) * two_pi;
    const float dim_t = powf(10000.0f, (float)(2 * (cc / 2)) / (float)feats);
    const float v = e / dim_t;
    out[idx] = (cc & 1) ? cosf(v) : sinf(v);
  }
}

// lo = x - trunc_tf32(x), the same expression the GEMM's splitter warps evaluate in shared memory
__global__ void __launch_bounds__(256)
tf32_residual_kernel(const float* __restrict__ x, float* __restrict__ lo, long long n) {
  pdl_wait();
  pdl_trigger();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float v = x[i];
    lo[i] = tf32_lo_exact(v);      // must equal the in-kernel split of itn_gemm_tf32 bit for bit
  }
}

// out[r, c] = (residual ? residual[r, c] : 0) + x[r, c] * keep(seed, site, row0 + r, c) / (1 - p); one thread per
// 4 consecutive columns (one Philox call).  In place (out == x) is fine.
__global__ void __launch_bounds__(256)
dropout_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ residual, long long ldr,
               float* __restrict__ out, long long ldo, long long rows, int cols, long long row0, DropParams d) {
  pdl_wait();
  pdl_trigger();
  const int groups = (cols + 3) >> 2;
  const long long total = rows * groups;
  const unsigned long long seed = *d.seed;
  const bool vec = ((cols | ldx | ldo | ldr) & 3) == 0 &&
                   ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) |
                     reinterpret_cast<uintptr_t>(residual)) & 15) == 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / groups;
    const int g = (int)(i - r * groups);
    const uint4 w = dropout_words(seed, d.site, (unsigned long long)(row0 + r), (unsigned int)g);
    const float k0 = w.x >= d.thr ? d.inv_keep : 0.f, k1 = w.y >= d.thr ? d.inv_keep : 0.f;
    const float k2 = w.z >= d.thr ? d.inv_keep : 0.f, k3 = w.w >= d.thr ? d.inv_keep : 0.f;
    const float* xp = x + r * ldx + 4 * g;
    float* op = out + r * ldo + 4 * g;
    const float* rp = residual ? residual + r * ldr + 4 * g : nullptr;
    if (vec) {
      const float4 v = *reinterpret_cast<const float4*>(xp);
      float4 y = make_float4(v.x * k0, v.y * k1, v.z * k2, v.w * k3);
      if (rp) {
        const float4 q = *reinterpret_cast<const float4*>(rp);
        y = make_float4(q.x + y.x, q.y + y.y, q.z + y.z, q.w + y.w);
      }
      *reinterpret_cast<float4*>(op) = y;
    } else {
      const float k[4] = {k0, k1, k2, k3};
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (4 * g + j < cols) op[j] = (rp ? rp[j] : 0.f) + xp[j] * k[j];
    }
  }
}

}  // namespace itn

using namespace itn;

extern "C" int itn_tf32_residual(const float* x, float* lo, long long n, void* stream) {
  ITN_REQUIRE(x && lo && n > 0, "tf32_residual: bad arguments");
  launch(tf32_residual_kernel, grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream), x, lo, n);
  return check_launch("tf32_residual_kernel");
}

extern "C" int itn_sgd_clip_update(const float* theta, long long theta_stride, const float* g,
                                   float* theta_out, float* theta_out_r, unsigned char* clip_mask,
                                   int groups,
                                   long long n, float lr, float clip, void* stream) {
  ITN_REQUIRE(theta && g && theta_out && groups > 0 && n > 0, "sgd_clip_update: bad arguments");
  ITN_REQUIRE(((uintptr_t)theta & 15) == 0 && ((uintptr_t)g & 15) == 0 &&
                  ((uintptr_t)theta_out & 15) == 0,
              "sgd_clip_update: pointers must be 16-byte aligned");
  ITN_REQUIRE(groups == 1 || ((n & 3) == 0 && (theta_stride & 3) == 0),
              "sgd_clip_update: n and theta_stride must be multiples of 4 when groups > 1");
  ITN_REQUIRE(!clip_mask || ((uintptr_t)clip_mask & 3) == 0, "sgd_clip_update: mask must be 4-byte aligned");
  ITN_REQUIRE(!theta_out_r || ((uintptr_t)theta_out_r & 15) == 0, "sgd_clip_update: theta_out_r must be 16-byte aligned");
  dim3 grid(grid_for(n >> 2, 256, 8), groups);
  launch(sgd_clip_update_kernel, grid, 256, 0, static_cast<cudaStream_t>(stream), 
      theta, theta_stride, g, theta_out, theta_out_r, clip_mask, n, lr, clip);
  return check_launch("sgd_clip_update_kernel");
}

extern "C" int itn_add(const float* a, const float* b, float* out, long long n, long long b_elems,
                       long long a_group, long long b_group_stride, int round_out, void* stream) {
  ITN_REQUIRE(a && b && out && n > 0 && b_elems > 0 && a_group > 0 && b_group_stride >= 0,
              "add: bad arguments");
  ITN_REQUIRE((((uintptr_t)a | (uintptr_t)b | (uintptr_t)out) & 15) == 0, "add: pointers must be 16-byte aligned");
  ITN_REQUIRE(((n | b_elems | a_group | b_group_stride) & 3) == 0,
              "add: n, b_elems, a_group, b_group_stride must be multiples of 4");
  ITN_REQUIRE(a_group % b_elems == 0, "add: a_group must be a multiple of b_elems");
  launch(add_kernel, grid_for(n >> 2, 256), 256, 0, static_cast<cudaStream_t>(stream), 
      a, b, out, n, b_elems, a_group, b_group_stride, round_out);
  return check_launch("add_kernel");
}

extern "C" int itn_copy2d(const float* src, long long lds, float* dst, long long ldd, long long rows,
                          int cols, int round_out, void* stream) {
  ITN_REQUIRE(src && dst && rows > 0 && cols > 0, "copy2d: bad arguments");
  launch(copy2d_kernel, grid_for(rows * cols, 256), 256, 0, static_cast<cudaStream_t>(stream), 
      src, lds, dst, ldd, rows, cols, round_out);
  return check_launch("copy2d_kernel");
}

extern "C" int itn_transpose(const float* src, float* dst, int groups, int rows, int cols,
                             long long src_group_stride, long long dst_group_stride, void* stream) {
  ITN_REQUIRE(src && dst && groups > 0 && rows > 0 && cols > 0, "transpose: bad arguments");
  ITN_REQUIRE(groups <= 65535, "transpose: too many groups");
  dim3 grid((cols + 31) / 32, (rows + 31) / 32, groups);
  launch(transpose_kernel, grid, 256, 0, static_cast<cudaStream_t>(stream), src, dst, rows, cols,
                                                                        src_group_stride, dst_group_stride);
  return check_launch("transpose_kernel");
}

extern "C" int itn_round_tf32(const float* src, float* dst, long long n, void* stream) {
  ITN_REQUIRE(src && dst && n > 0, "round_tf32: bad arguments");
  ITN_REQUIRE((((uintptr_t)src | (uintptr_t)dst) & 15) == 0, "round_tf32: pointers must be 16-byte aligned");
  launch(round_tf32_kernel, grid_for(n >> 2, 256), 256, 0, static_cast<cudaStream_t>(stream), src, dst, n);
  return check_launch("round_tf32_kernel");
}

extern "C" int itn_sigmoid_fwd(const float* x, float* y, long long n, void* stream) {
  ITN_REQUIRE(x && y && n > 0, "sigmoid_fwd: bad arguments");
  launch(sigmoid_fwd_kernel, grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream), x, y, n);
  return check_launch("sigmoid_fwd_kernel");
}

extern "C" int itn_sigmoid_bwd(const float* dy, const float* y, float* dx, long long n,
                               void* stream) {
  ITN_REQUIRE(dy && y && dx && n > 0, "sigmoid_bwd: bad arguments");
  launch(sigmoid_bwd_kernel, grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream), dy, y, dx, n);
  return check_launch("sigmoid_bwd_kernel");
}

extern "C" int itn_l2norm_fwd_bwd(const float* x, float* loss, float* dx, int groups, int n,
                                  void* stream) {
  ITN_REQUIRE(x && groups > 0 && n > 0, "l2norm_fwd_bwd: bad arguments");
  launch(l2norm_fwd_bwd_kernel, groups, 256, 0, static_cast<cudaStream_t>(stream), x, loss, dx, n);
  return check_launch("l2norm_fwd_bwd_kernel");
}

extern "C" int itn_pos_embed_sine(const unsigned char* mask, float* pos, int frames, int h, int w,
                                  int feats, void* stream) {
  ITN_REQUIRE(mask && pos && frames > 0 && h > 0 && w > 0 && feats > 0, "pos_embed_sine: bad arguments");
  const size_t smem = 2ull * h * w * sizeof(float);
  ITN_REQUIRE(smem <= 48 * 1024, "pos_embed_sine: feature map %dx%d too large", h, w);
  launch(pos_embed_sine_kernel, frames, 256, smem, static_cast<cudaStream_t>(stream), mask, pos, h, w, feats);
  return check_launch("pos_embed_sine_kernel");
}

extern "C" int itn_dropout(const float* x, long long ldx, const float* residual, long long ldr, float* out,
                           long long ldo, long long rows, int cols, long long row0, float p,
                           const unsigned long long* seed, unsigned int site, void* stream) {
  ITN_REQUIRE(x && out && seed && rows > 0 && cols > 0, "dropout: bad arguments");
  ITN_REQUIRE(p >= 0.f && p < 1.f, "dropout: p must be in [0, 1)");
  itn::DropParams d;
  d.seed = seed;
  d.site = site;
  d.thr = (unsigned int)((double)p * 4294967296.0);
  d.inv_keep = 1.0f / (1.0f - p);
  const long long total = rows * ((cols + 3) / 4);
  launch(dropout_kernel, grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream), x, ldx, residual, ldr, out,
         ldo, rows, cols, row0, d);
  return check_launch("dropout_kernel");
}
