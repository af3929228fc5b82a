// HungarianMatcher cost matrix (block-diagonal only).  One warp per (frame,
// query): warp-shuffle softmax statistics over the class logits, then one lane
// per target for the class / L1 / GIoU terms.  Contract in interactron_b200.h.
#include "itn_common.cuh"

namespace itn {

__global__ void __launch_bounds__(256)
matcher_cost_kernel(const float* __restrict__ logits, const float* __restrict__ boxes,
                    const float* __restrict__ tgt_boxes, const long long* __restrict__ tgt_labels,
                    const int* __restrict__ tgt_off, float* __restrict__ cost, int frames,
                    int queries, int classes, float w_class, float w_bbox, float w_giou) {
  pdl_wait();
  pdl_trigger();
  const int wq = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (wq >= frames * queries) return;
  const int lane = threadIdx.x & 31;
  const int f = wq / queries, q = wq - f * queries;
  const float* lg = logits + (long long)wq * classes;
  float m = -INFINITY;
  for (int c = lane; c < classes; c += 32) m = fmaxf(m, lg[c]);
  m = warp_max(m);
  float s = 0.f;
  for (int c = lane; c < classes; c += 32) s += expf(lg[c] - m);
  s = warp_sum(s);
  const float4 b = reinterpret_cast<const float4*>(boxes)[wq];  // cx, cy, w, h
  const float bx0 = b.x - 0.5f * b.z, by0 = b.y - 0.5f * b.w;
  const float bx1 = b.x + 0.5f * b.z, by1 = b.y + 0.5f * b.w;
  const float area_b = (bx1 - bx0) * (by1 - by0);
  const int t0 = tgt_off[f], t1 = tgt_off[f + 1];
  const int nt = t1 - t0;
  // frame f's block starts after the blocks of the previous frames
  float* out = cost + (long long)queries * t0 + (long long)q * nt;
  for (int t = lane; t < nt; t += 32) {
    const float4 g = reinterpret_cast<const float4*>(tgt_boxes)[t0 + t];
    const float prob = expf(lg[tgt_labels[t0 + t]] - m) / s;
    const float l1 = fabsf(b.x - g.x) + fabsf(b.y - g.y) + fabsf(b.z - g.z) + fabsf(b.w - g.w);
    const float gx0 = g.x - 0.5f * g.z, gy0 = g.y - 0.5f * g.w;
    const float gx1 = g.x + 0.5f * g.z, gy1 = g.y + 0.5f * g.w;
    const float area_g = (gx1 - gx0) * (gy1 - gy0);
    const float iw = fmaxf(fminf(bx1, gx1) - fmaxf(bx0, gx0), 0.f);
    const float ih = fmaxf(fminf(by1, gy1) - fmaxf(by0, gy0), 0.f);
    const float inter = iw * ih;
    const float uni = area_b + area_g - inter;
    const float iou = inter / uni;
    const float cw = fmaxf(fmaxf(bx1, gx1) - fminf(bx0, gx0), 0.f);
    const float ch = fmaxf(fmaxf(by1, gy1) - fminf(by0, gy0), 0.f);
    const float carea = cw * ch;
    const float giou = iou - (carea - uni) / carea;
    out[t] = w_bbox * l1 + w_class * (-prob) + w_giou * (-giou);
  }
}

}  // namespace itn

extern "C" int itn_matcher_cost(const float* logits, const float* boxes, const float* tgt_boxes,
                                const long long* tgt_labels, const int* tgt_off, float* cost,
                                int frames, int queries, int classes, float w_class, float w_bbox,
                                float w_giou, void* stream) {
  ITN_REQUIRE(logits && boxes && tgt_boxes && tgt_labels && tgt_off && cost,
              "matcher_cost: null pointer");
  ITN_REQUIRE(frames > 0 && queries > 0 && classes > 0, "matcher_cost: bad sizes");
  ITN_REQUIRE((((uintptr_t)boxes | (uintptr_t)tgt_boxes) & 15) == 0,
              "matcher_cost: boxes must be 16-byte aligned");
  const int warps = frames * queries;
  itn::launch(itn::matcher_cost_kernel, (warps + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream), 
      logits, boxes, tgt_boxes, tgt_labels, tgt_off, cost, frames, queries, classes, w_class,
      w_bbox, w_giou);
  return itn::check_launch("matcher_cost_kernel");
}
