// SetCriterion (reference models/detr_models/detr.py:111-265) on the device: weighted cross-entropy
// over the class logits, L1 + GIoU on the matched boxes, the two logging metrics, and - optionally -
// the gradient of  w_ce*loss_ce + w_bbox*loss_bbox + w_giou*loss_giou  wrt logits and boxes (what
// `.backward()` hands to the detector at models/interactron.py:121-123,133).  The Hungarian
// assignment is an input (itn_matcher_cost + host LSAP).  A "group" is one criterion call of the
// reference (the F frames of one episode); every group is normalised by its own num_boxes.
// All reductions run in a fixed order: results are deterministic.  Contract in interactron_b200.h.
#include "itn_common.cuh"

namespace itn {

// rows without a match are "no object" (class index classes-1), detr.py:120-122
__global__ void __launch_bounds__(256)
criterion_fill_targets_kernel(int* __restrict__ tgt_class, int rows, int no_object) {
  pdl_wait();
  pdl_trigger();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < rows) tgt_class[r] = no_object;
}

__global__ void __launch_bounds__(256)
criterion_scatter_targets_kernel(int* __restrict__ tgt_class, const int* __restrict__ match_row,
                                 const int* __restrict__ match_tgt,
                                 const long long* __restrict__ tgt_labels, int n_match) {
  pdl_wait();
  pdl_trigger();
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m < n_match) tgt_class[match_row[m]] = (int)tgt_labels[match_tgt[m]];
}

// One warp per (frame, query) row: log-sum-exp statistics, weighted NLL, arg-max.
// row_stats[r] = {w * nll, w, max, log(sum exp(l - max))};  row_argmax[r] = first arg-max.
__global__ void __launch_bounds__(256)
criterion_rows_kernel(const float* __restrict__ logits, const int* __restrict__ tgt_class,
                      float4* __restrict__ row_stats, int* __restrict__ row_argmax, int rows, int classes,
                      float background_c) {
  pdl_wait();
  pdl_trigger();
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* lg = logits + (long long)r * classes;
  float m = -INFINITY;
  int am = 0x7fffffff;
  for (int c = lane; c < classes; c += 32) {
    const float v = lg[c];
    if (v > m) {
      m = v;
      am = c;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, o);
    const int oa = __shfl_xor_sync(0xffffffffu, am, o);
    if (om > m || (om == m && oa < am)) {
      m = om;
      am = oa;
    }
  }
  float s = 0.f;
  for (int c = lane; c < classes; c += 32) s += expf(lg[c] - m);
  s = warp_sum(s);
  if (lane == 0) {
    const int tc = tgt_class[r];
    const float w = tc == classes - 1 ? background_c : 1.0f;
    const float ls = logf(s);
    row_stats[r] = make_float4(w * (ls + m - lg[tc]), w, m, ls);
    row_argmax[r] = am;
  }
}

struct GiouTerms {
  float giou;
  float g[4];   // d giou / d (x0, y0, x1, y1) of the source box
};

// GIoU of two xyxy boxes and its gradient wrt the first (util/box_ops.py:38-58 + autograd).
__device__ __forceinline__ GiouTerms giou_with_grad(float sx0, float sy0, float sx1, float sy1, float tx0,
                                                    float ty0, float tx1, float ty1) {
  const float sw = sx1 - sx0, sh = sy1 - sy0;
  const float area_s = sw * sh, area_t = (tx1 - tx0) * (ty1 - ty0);
  const float iw_raw = fminf(sx1, tx1) - fmaxf(sx0, tx0), ih_raw = fminf(sy1, ty1) - fmaxf(sy0, ty0);
  const float iw = fmaxf(iw_raw, 0.f), ih = fmaxf(ih_raw, 0.f);
  const float inter = iw * ih;
  const float uni = area_s + area_t - inter;
  const float iou = inter / uni;
  const float cw = fmaxf(fmaxf(sx1, tx1) - fminf(sx0, tx0), 0.f);
  const float ch = fmaxf(fmaxf(sy1, ty1) - fminf(sy0, ty0), 0.f);
  const float areac = cw * ch;
  GiouTerms o;
  o.giou = iou - (areac - uni) / areac;
  // d giou = A d inter + B d area_s + C d areac
  const float A = 1.0f / uni + inter / (uni * uni) - 1.0f / areac;
  const float B = -inter / (uni * uni) + 1.0f / areac;
  const float C = -uni / (areac * areac);
  const float iw_on = iw_raw >= 0.f ? 1.f : 0.f, ih_on = ih_raw >= 0.f ? 1.f : 0.f;
  // x0
  o.g[0] = A * (-(sx0 > tx0 ? 1.f : 0.f) * iw_on * ih) + B * (-sh) + C * (-(sx0 < tx0 ? 1.f : 0.f) * ch);
  o.g[1] = A * (-(sy0 > ty0 ? 1.f : 0.f) * ih_on * iw) + B * (-sw) + C * (-(sy0 < ty0 ? 1.f : 0.f) * cw);
  o.g[2] = A * ((sx1 < tx1 ? 1.f : 0.f) * iw_on * ih) + B * sh + C * ((sx1 > tx1 ? 1.f : 0.f) * ch);
  o.g[3] = A * ((sy1 < ty1 ? 1.f : 0.f) * ih_on * iw) + B * sw + C * ((sy1 > ty1 ? 1.f : 0.f) * cw);
  return o;
}

// One thread per matched (prediction, target) pair.
// match_stats[m] = {sum |src - tgt|, 1 - giou, top-1 correct};  dboxes rows of matched predictions.
__global__ void __launch_bounds__(256)
criterion_match_kernel(const float* __restrict__ boxes, const float* __restrict__ tgt_boxes,
                       const long long* __restrict__ tgt_labels, const int* __restrict__ tgt_off,
                       const int* __restrict__ match_row, const int* __restrict__ match_tgt,
                       const int* __restrict__ row_argmax, float* __restrict__ match_stats,
                       float* __restrict__ dboxes, int n_match, int rows_per_group, int frames_per_group,
                       float w_bbox, float w_giou) {
  pdl_wait();
  pdl_trigger();
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n_match) return;
  const int r = match_row[m], t = match_tgt[m];
  const float4 s = reinterpret_cast<const float4*>(boxes)[r];
  const float4 g = reinterpret_cast<const float4*>(tgt_boxes)[t];
  const float l1 = fabsf(s.x - g.x) + fabsf(s.y - g.y) + fabsf(s.z - g.z) + fabsf(s.w - g.w);
  const GiouTerms gi = giou_with_grad(s.x - 0.5f * s.z, s.y - 0.5f * s.w, s.x + 0.5f * s.z, s.y + 0.5f * s.w,
                                      g.x - 0.5f * g.z, g.y - 0.5f * g.w, g.x + 0.5f * g.z, g.y + 0.5f * g.w);
  match_stats[3 * m + 0] = l1;
  match_stats[3 * m + 1] = 1.0f - gi.giou;
  match_stats[3 * m + 2] = row_argmax[r] == (int)tgt_labels[t] ? 1.0f : 0.0f;
  if (dboxes) {
    const int grp = r / rows_per_group;
    const int nt = tgt_off[(grp + 1) * frames_per_group] - tgt_off[grp * frames_per_group];
    const float inv_nb = 1.0f / (float)(nt > 1 ? nt : 1);
    auto sgn = [](float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); };
    // cxcywh <- xyxy chain: d/dcx = g0 + g2, d/dw = (g2 - g0) / 2 (same for y / h); loss has -giou
    float4 d;
    d.x = (w_bbox * sgn(s.x - g.x) - w_giou * (gi.g[0] + gi.g[2])) * inv_nb;
    d.y = (w_bbox * sgn(s.y - g.y) - w_giou * (gi.g[1] + gi.g[3])) * inv_nb;
    d.z = (w_bbox * sgn(s.z - g.z) - w_giou * 0.5f * (gi.g[2] - gi.g[0])) * inv_nb;
    d.w = (w_bbox * sgn(s.w - g.w) - w_giou * 0.5f * (gi.g[3] - gi.g[1])) * inv_nb;
    reinterpret_cast<float4*>(dboxes)[r] = d;
  }
}

// One block per group: fixed-order tree reductions -> losses[g] = {loss_ce, class_error,
// cardinality_error, loss_bbox, loss_giou};  group_wsum[g] = sum of the class weights (for the gradient).
__global__ void __launch_bounds__(256)
criterion_finalize_kernel(const float4* __restrict__ row_stats, const int* __restrict__ row_argmax,
                          const float* __restrict__ match_stats, const int* __restrict__ match_off,
                          const int* __restrict__ tgt_off, float* __restrict__ losses,
                          float* __restrict__ group_wsum, int frames_per_group, int queries, int classes) {
  pdl_wait();
  pdl_trigger();
  __shared__ float red[6][256];
  __shared__ float card[256];
  const int g = blockIdx.x, tid = threadIdx.x;
  const int rows = frames_per_group * queries;
  float a_nll = 0.f, a_w = 0.f, a_l1 = 0.f, a_gi = 0.f, a_ok = 0.f;
  for (int r = tid; r < rows; r += 256) {
    const float4 st = row_stats[g * rows + r];
    a_nll += st.x;
    a_w += st.y;
  }
  const int m0 = match_off[g], m1 = match_off[g + 1];
  for (int m = m0 + tid; m < m1; m += 256) {
    a_l1 += match_stats[3 * m + 0];
    a_gi += match_stats[3 * m + 1];
    a_ok += match_stats[3 * m + 2];
  }
  // cardinality: |#(argmax != no-object) - #targets| per frame, averaged over the frames
  float a_card = 0.f;
  for (int f = tid; f < frames_per_group; f += 256) {
    int cnt = 0;
    for (int q = 0; q < queries; ++q) cnt += row_argmax[(g * frames_per_group + f) * queries + q] != classes - 1;
    const int nt = tgt_off[g * frames_per_group + f + 1] - tgt_off[g * frames_per_group + f];
    a_card += fabsf((float)cnt - (float)nt);
  }
  red[0][tid] = a_nll; red[1][tid] = a_w; red[2][tid] = a_l1; red[3][tid] = a_gi; red[4][tid] = a_ok;
  card[tid] = a_card;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) {
#pragma unroll
      for (int k = 0; k < 5; ++k) red[k][tid] += red[k][tid + o];
      card[tid] += card[tid + o];
    }
    __syncthreads();
  }
  if (tid == 0) {
    const int nt = tgt_off[(g + 1) * frames_per_group] - tgt_off[g * frames_per_group];
    const float nb = (float)(nt > 1 ? nt : 1);
    const int nm = m1 - m0;
    losses[5 * g + 0] = red[0][0] / red[1][0];
    losses[5 * g + 1] = nm > 0 ? 100.0f - red[4][0] * (100.0f / (float)nm) : 100.0f;
    losses[5 * g + 2] = card[0] / (float)frames_per_group;
    losses[5 * g + 3] = red[2][0] / nb;
    losses[5 * g + 4] = red[3][0] / nb;
    group_wsum[g] = red[1][0];
  }
}

// dlogits[r, c] = w_ce * w_r / W_group * (softmax(l_r)[c] - [c == target_r])
__global__ void __launch_bounds__(256)
criterion_dlogits_kernel(const float* __restrict__ logits, const int* __restrict__ tgt_class,
                         const float4* __restrict__ row_stats, const float* __restrict__ group_wsum,
                         float* __restrict__ dlogits, int rows, int rows_per_group, int classes, float w_ce) {
  pdl_wait();
  pdl_trigger();
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const float4 st = row_stats[r];
  const float coef = w_ce * st.y / group_wsum[r / rows_per_group];
  const int tc = tgt_class[r];
  const float* lg = logits + (long long)r * classes;
  float* dl = dlogits + (long long)r * classes;
  for (int c = lane; c < classes; c += 32)
    dl[c] = coef * (expf(lg[c] - st.z - st.w) - (c == tc ? 1.0f : 0.0f));
}

__global__ void __launch_bounds__(256)
zero_kernel(float* __restrict__ p, long long n) {
  pdl_wait();
  pdl_trigger();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = 0.f;
}

}  // namespace itn

using namespace itn;

extern "C" long long itn_criterion_scratch_bytes(int rows, int n_match, int groups) {
  // tgt_class[rows] int | row_argmax[rows] int | row_stats[rows] float4 | match_stats[3*n_match] | wsum[groups]
  long long b = (long long)rows * 4 * 2;
  b = (b + 15) / 16 * 16;
  b += (long long)rows * 16 + (long long)(n_match > 0 ? n_match : 1) * 12 + (long long)groups * 4;
  return (b + 15) / 16 * 16;
}

extern "C" int itn_criterion(const float* logits, const float* boxes, const float* tgt_boxes,
                             const long long* tgt_labels, const int* tgt_off, const int* match_row,
                             const int* match_tgt, const int* match_off, int n_match, int groups,
                             int frames_per_group, int queries, int classes, float background_c, float w_ce,
                             float w_bbox, float w_giou, float* losses, float* dlogits, float* dboxes,
                             void* scratch, void* stream) {
  ITN_REQUIRE(logits && boxes && tgt_off && match_off && losses && scratch, "criterion: null pointer");
  ITN_REQUIRE(groups > 0 && frames_per_group > 0 && queries > 0 && classes > 1, "criterion: bad sizes");
  ITN_REQUIRE(n_match >= 0 && (n_match == 0 || (tgt_boxes && tgt_labels && match_row && match_tgt)),
              "criterion: matches given without targets");
  ITN_REQUIRE((((uintptr_t)boxes | (uintptr_t)tgt_boxes | (uintptr_t)dboxes | (uintptr_t)scratch) & 15) == 0,
              "criterion: boxes / dboxes / scratch must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int rows = groups * frames_per_group * queries;
  const int rpg = frames_per_group * queries;
  char* base = static_cast<char*>(scratch);
  int* tgt_class = reinterpret_cast<int*>(base);
  int* row_argmax = tgt_class + rows;
  long long off = ((long long)rows * 8 + 15) / 16 * 16;
  float4* row_stats = reinterpret_cast<float4*>(base + off);
  off += (long long)rows * 16;
  float* match_stats = reinterpret_cast<float*>(base + off);
  off += (long long)(n_match > 0 ? n_match : 1) * 12;
  float* wsum = reinterpret_cast<float*>(base + off);

  launch(criterion_fill_targets_kernel, (rows + 255) / 256, 256, 0, s, tgt_class, rows, classes - 1);
  int rc = check_launch("criterion_fill_targets_kernel");
  if (rc) return rc;
  if (n_match > 0) {
    launch(criterion_scatter_targets_kernel, (n_match + 255) / 256, 256, 0, s, tgt_class, match_row, match_tgt,
           tgt_labels, n_match);
    if ((rc = check_launch("criterion_scatter_targets_kernel"))) return rc;
  }
  launch(criterion_rows_kernel, (rows + 7) / 8, 256, 0, s, logits, (const int*)tgt_class, row_stats, row_argmax,
         rows, classes, background_c);
  if ((rc = check_launch("criterion_rows_kernel"))) return rc;
  if (dboxes) {
    launch(zero_kernel, (unsigned)((rows * 4LL + 255) / 256), 256, 0, s, dboxes, rows * 4LL);
    if ((rc = check_launch("zero_kernel"))) return rc;
  }
  if (n_match > 0) {
    launch(criterion_match_kernel, (n_match + 255) / 256, 256, 0, s, boxes, tgt_boxes, tgt_labels, tgt_off,
           match_row, match_tgt, (const int*)row_argmax, match_stats, dboxes, n_match, rpg, frames_per_group,
           w_bbox, w_giou);
    if ((rc = check_launch("criterion_match_kernel"))) return rc;
  }
  launch(criterion_finalize_kernel, groups, 256, 0, s, (const float4*)row_stats, (const int*)row_argmax,
         (const float*)match_stats, match_off, tgt_off, losses, wsum, frames_per_group, queries, classes);
  if ((rc = check_launch("criterion_finalize_kernel"))) return rc;
  if (dlogits) {
    launch(criterion_dlogits_kernel, (rows + 7) / 8, 256, 0, s, logits, (const int*)tgt_class,
           (const float4*)row_stats, (const float*)wsum, dlogits, rows, rpg, classes, w_ce);
    rc = check_launch("criterion_dlogits_kernel");
  }
  return rc;
}
