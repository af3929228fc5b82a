// Evaluator post-processing of predict()'s output, one CTA per image:
//   scores, cats = logits.softmax(-1).max(-1)            engine/random_policy_evaluator.py:66
//   drop predictions whose argmax is the background class  :70-73
//   boxes cxcywh -> xyxy                                   :65 (detr_models/util/box_ops.py:8-12)
//   torchvision.ops.nms(boxes, scores, iou_threshold)      :75   (class-agnostic)
// The reference does this per image with ~10 device round trips; here one launch handles every image
// of a predict() batch and the host reads four small arrays back once.
//   phase 1  warp per query: max / first arg-max / sum of exp over the classes (shuffles);
//            the top probability is exp(0)/sum = 1/sum, as in softmax-then-max
//   phase 2  rank sort of the kept queries by decreasing score (ties: lower query index first,
//            the order of a stable descending sort), then the greedy suppression loop of torchvision's
//            nms kernel, one candidate per iteration with the overlaps of all later boxes in parallel
// IoU arithmetic is written in torchvision's operation order without fused multiply-adds, so the
// keep / suppress decisions are bit-identical to torchvision.ops.nms on the same boxes and scores.
#include "itn_common.cuh"

namespace itn {

constexpr int kMaxQ = 128;

__global__ void __launch_bounds__(256)
detect_postprocess_kernel(const float* __restrict__ logits, const float* __restrict__ boxes, int Q, int C,
                          int background, float iou_thr, int* __restrict__ count, int* __restrict__ keep_idx,
                          float* __restrict__ score, int* __restrict__ cat, float* __restrict__ xyxy) {
  pdl_wait();
  pdl_trigger();
  __shared__ float s_score[kMaxQ];
  __shared__ int s_cat[kMaxQ];
  __shared__ float4 s_box[kMaxQ];
  __shared__ float s_area[kMaxQ];
  __shared__ int s_order[kMaxQ];      // sorted position -> query
  __shared__ int s_dead[kMaxQ];       // by sorted position
  __shared__ int s_n;
  const int img = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* lg_img = logits + (long long)img * Q * C;
  if (threadIdx.x == 0) s_n = 0;
  for (int q = warp; q < Q; q += 8) {
    const float* lg = lg_img + (long long)q * C;
    float m = -INFINITY;
    int am = 0x7fffffff;
    for (int c = lane; c < C; c += 32) {
      const float v = lg[c];
      if (v > m) { m = v; am = c; }          // strict: the first maximum of this lane's slice
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float m2 = __shfl_xor_sync(0xffffffffu, m, o);
      const int a2 = __shfl_xor_sync(0xffffffffu, am, o);
      if (m2 > m || (m2 == m && a2 < am)) { m = m2; am = a2; }
    }
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += expf(lg[c] - m);
    s = warp_sum(s);
    if (lane == 0) {
      s_score[q] = 1.0f / s;
      s_cat[q] = am;
      const float4 b = reinterpret_cast<const float4*>(boxes)[(long long)img * Q + q];   // cx, cy, w, h
      float4 r;
      r.x = b.x - 0.5f * b.z; r.y = b.y - 0.5f * b.w;      // 0.5 * w is exact: contraction cannot change these
      r.z = b.x + 0.5f * b.z; r.w = b.y + 0.5f * b.w;
      s_box[q] = r;
      s_area[q] = __fmul_rn(r.z - r.x, r.w - r.y);
    }
  }
  __syncthreads();
  // rank of every kept query among the kept ones
  const int q = threadIdx.x;
  const bool live = q < Q && s_cat[q] != background;
  if (live) {
    const float sq = s_score[q];
    int r = 0;
    for (int j = 0; j < Q; ++j)
      if (s_cat[j] != background && (s_score[j] > sq || (s_score[j] == sq && j < q))) ++r;
    s_order[r] = q;
    s_dead[r] = 0;
    atomicAdd(&s_n, 1);
  }
  __syncthreads();
  const int n = s_n;
  // greedy suppression in score order
  for (int i = 0; i < n; ++i) {
    if (!s_dead[i]) {                        // uniform: every thread reads the same flag
      const int j = threadIdx.x;
      if (j > i && j < n && !s_dead[j]) {
        const float4 a = s_box[s_order[i]], b = s_box[s_order[j]];
        const float w = fmaxf(0.f, fminf(a.z, b.z) - fmaxf(a.x, b.x));
        const float h = fmaxf(0.f, fminf(a.w, b.w) - fmaxf(a.y, b.y));
        const float inter = __fmul_rn(w, h);
        const float ovr = inter / (s_area[s_order[i]] + s_area[s_order[j]] - inter);
        if (ovr > iou_thr) s_dead[j] = 1;
      }
    }
    __syncthreads();
  }
  // compact the survivors (still in score order)
  if (threadIdx.x == 0) {
    int k = 0;
    for (int i = 0; i < n; ++i) {
      if (s_dead[i]) continue;
      const int qq = s_order[i];
      const long long o = (long long)img * Q + k;
      keep_idx[o] = qq;
      score[o] = s_score[qq];
      cat[o] = s_cat[qq];
      reinterpret_cast<float4*>(xyxy)[o] = s_box[qq];
      ++k;
    }
    count[img] = k;
    for (; k < Q; ++k) keep_idx[(long long)img * Q + k] = -1;
  }
}

}  // namespace itn

extern "C" int itn_detect_postprocess(const float* logits, const float* boxes, int images, int queries,
                                      int classes, int background_class, float iou_threshold, int* count,
                                      int* keep_idx, float* score, int* cat, float* xyxy, void* stream) {
  ITN_REQUIRE(logits && boxes && count && keep_idx && score && cat && xyxy, "detect_postprocess: null pointer");
  ITN_REQUIRE(images > 0 && classes > 0, "detect_postprocess: bad sizes");
  ITN_REQUIRE(queries > 0 && queries <= itn::kMaxQ, "detect_postprocess: 1..128 queries per image");
  ITN_REQUIRE((((uintptr_t)boxes | (uintptr_t)xyxy) & 15) == 0, "detect_postprocess: boxes must be 16-byte aligned");
  itn::launch(itn::detect_postprocess_kernel, dim3(images), 256, 0, static_cast<cudaStream_t>(stream), logits,
              boxes, queries, classes, background_class, iou_threshold, count, keep_idx, score, cat, xyxy);
  return itn::check_launch("detect_postprocess_kernel");
}
