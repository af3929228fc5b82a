// Channels-last convolution support for the frozen ResNet trunk: im2col gather (any kernel size /
// stride / padding / dilation) feeding the tf32x3 tensor-core GEMM, and 3x3/2 max-pooling.
// Both are pure data movement: coalesced 128-bit accesses along the channel dim.
#include "itn_common.cuh"

namespace itn {

// dst[(n,ho,wo)][(ky*kw+kx)*C + c] = src[n][ho*s - p + ky*d][wo*s - p + kx*d][c]  (0 outside).
// One warp per output pixel: the pixel is decoded once (32-bit divisions), then the lanes sweep the
// row's kh*kw*C/VEC vectors 32 at a time, i.e. 512 contiguous bytes per store instruction.
// (The first version decoded every vector with four 64-bit divisions and was ALU-bound at
// ~2.1 TB/s of output; this one is bound by the HBM write.)
template <int VEC>
__global__ void __launch_bounds__(256)
im2col_nhwc_kernel(const float* __restrict__ src, float* __restrict__ dst, int N, int H, int W, int C,
                   int kh, int kw, int stride, int pad, int dil, int Ho, int Wo, long long ld,
                   int cv_shift) {
  pdl_wait();
  pdl_trigger();
  const int cv = C / VEC;                              // vectors per tap
  const int per_row = kh * kw * cv;
  const int rows = N * Ho * Wo;
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < rows; row += warps) {
    const int wo = row % Wo;
    const int t = row / Wo;
    const int ho = t % Ho;
    const int n = t / Ho;
    const int h0 = ho * stride - pad, w0 = wo * stride - pad;
    const float* sn = src + (long long)n * H * W * C;
    float* d = dst + (long long)row * ld;
    for (int e = lane; e < per_row; e += 32) {
      const int tap = cv_shift >= 0 ? e >> cv_shift : e / cv;
      const int c = (e - tap * cv) * VEC;
      const int ky = tap / kw, kx = tap - ky * kw;
      const int hi = h0 + ky * dil, wi = w0 + kx * dil;
      const bool in = hi >= 0 && hi < H && wi >= 0 && wi < W;
      const float* sp = sn + ((long long)hi * W + wi) * C + c;
      if (VEC == 4) {
        reinterpret_cast<float4*>(d)[e] = in ? *reinterpret_cast<const float4*>(sp) : make_float4(0.f, 0.f, 0.f, 0.f);
      } else {
        d[e] = in ? *sp : 0.f;
      }
    }
  }
}

// Zero the padding columns [kcols, ld) of an im2col matrix (only the 7x7x3 stem has any).
__global__ void __launch_bounds__(256)
zero_cols_kernel(float* __restrict__ dst, long long rows, int c0, int c1, long long ld) {
  pdl_wait();
  pdl_trigger();
  const int w = c1 - c0;
  const long long total = rows * w;
  const long long step = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += step)
    dst[(i / w) * ld + c0 + (int)(i % w)] = 0.f;
}

__global__ void __launch_bounds__(256)
maxpool3x3s2_nhwc_kernel(const float* __restrict__ src, float* __restrict__ dst, int N, int H, int W,
                         int C, int Ho, int Wo) {
  pdl_wait();
  pdl_trigger();
  const int cv = C / 4;
  const long long total = (long long)N * Ho * Wo * cv;
  const long long step = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += step) {
    const int c = (int)(i % cv) * 4;
    long long t = i / cv;
    const int wo = (int)(t % Wo);
    t /= Wo;
    const int ho = (int)(t % Ho);
    const int n = (int)(t / Ho);
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int hi = ho * 2 - 1 + ky;
      if (hi < 0 || hi >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int wi = wo * 2 - 1 + kx;
        if (wi < 0 || wi >= W) continue;
        const float4 v = *reinterpret_cast<const float4*>(src + (((long long)n * H + hi) * W + wi) * C + c);
        m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
      }
    }
    *reinterpret_cast<float4*>(dst + (((long long)n * Ho + ho) * Wo + wo) * C + c) = m;
  }
}

static inline unsigned conv_grid(long long items) {
  long long b = (items + 255) / 256;
  const long long cap = 148LL * 16;
  return (unsigned)(b > cap ? cap : (b < 1 ? 1 : b));
}

}  // namespace itn

using namespace itn;

extern "C" int itn_im2col_nhwc(const float* src, float* dst, int N, int H, int W, int C, int kh, int kw,
                               int stride, int pad, int dil, int Ho, int Wo, long long ld, void* stream) {
  ITN_REQUIRE(src && dst && N > 0 && H > 0 && W > 0 && C > 0 && kh > 0 && kw > 0 && stride > 0 && dil > 0 &&
                  Ho > 0 && Wo > 0, "im2col_nhwc: bad arguments");
  const long long kcols = (long long)kh * kw * C;
  ITN_REQUIRE(ld >= kcols, "im2col_nhwc: ld (%lld) < kh*kw*C (%lld)", ld, kcols);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool vec = (C % 4 == 0) && (ld % 4 == 0) && ((((uintptr_t)src | (uintptr_t)dst) & 15) == 0);
  const long long rows = (long long)N * Ho * Wo;
  ITN_REQUIRE(rows < (1LL << 31) / 32 && kcols < (1LL << 31), "im2col_nhwc: problem too large for 32-bit indexing");
  const int cvv = vec ? C / 4 : C;
  int cv_shift = -1;
  for (int b = 0; b < 31; ++b)
    if ((1 << b) == cvv) cv_shift = b;
  const unsigned grid = conv_grid(rows * 32);
  if (vec) {
    launch(im2col_nhwc_kernel<4>, grid, 256, 0, s, src, dst, N, H, W, C, kh, kw, stride, pad, dil, Ho, Wo, ld, cv_shift);
  } else {
    launch(im2col_nhwc_kernel<1>, grid, 256, 0, s, src, dst, N, H, W, C, kh, kw, stride, pad, dil, Ho, Wo, ld, cv_shift);
  }
  int rc = check_launch("im2col_nhwc_kernel");
  if (rc) return rc;
  if (ld > kcols) {
    launch(zero_cols_kernel, conv_grid(rows * (ld - kcols)), 256, 0, s, dst, rows, (int)kcols, (int)ld, ld);
    rc = check_launch("zero_cols_kernel");
  }
  return rc;
}

extern "C" int itn_maxpool3x3s2_nhwc(const float* src, float* dst, int N, int H, int W, int C, int Ho, int Wo,
                                     void* stream) {
  ITN_REQUIRE(src && dst && N > 0 && H > 0 && W > 0 && C > 0 && (C % 4) == 0, "maxpool3x3s2_nhwc: bad arguments");
  ITN_REQUIRE(Ho == (H + 2 - 3) / 2 + 1 && Wo == (W + 2 - 3) / 2 + 1, "maxpool3x3s2_nhwc: bad output size");
  launch(maxpool3x3s2_nhwc_kernel, conv_grid((long long)N * Ho * Wo * (C / 4)), 256, 0, static_cast<cudaStream_t>(stream), src, dst, N, H, W, C, Ho, Wo);
  return check_launch("maxpool3x3s2_nhwc_kernel");
}
