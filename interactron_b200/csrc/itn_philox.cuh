// Counter-based dropout masks (Philox4x32, 7 rounds).  The reference trainers run the models in train() mode
// (engine/interactron_trainer.py:73), where nn.Dropout(p=0.1) sits after every attention softmax, sub-layer
// output and FFN activation (models/detr_models/transformer.py:154-159,219-230; models/gpt.py:51,56,72,195).
// PyTorch's own generator cannot be reproduced, so the mask of an element is a pure function of
//   (seed, site, row, column)      site = index of the dropout call in forward order of one pass
// which lets the backward pass, the dual-number (second-order) pass and the fused attention kernels regenerate
// it instead of storing it.  One Philox call yields the keep-bits of 4 consecutive columns:
//   ctr = {row_lo, row_hi, column / 4, site}   key = {seed_lo, seed_hi}   keep[c % 4] = word[c % 4] >= p * 2^32.
// oracle/philox.py restates this in numpy; tests compare the two bit for bit.
#pragma once
#include <cstdint>

namespace itn {

struct DropParams {
  const unsigned long long* seed;   // device pointer (graph-replay safe: the host rewrites it between steps); null = off
  unsigned int site;
  unsigned int thr;                 // p * 2^32 (drop when word < thr)
  float inv_keep;                   // 1 / (1 - p)
};

__device__ __forceinline__ uint4 philox4x32_7(uint4 c, uint2 k) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 7; ++r) {
    const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0;
    k.y += W1;
  }
  return c;
}

// keep-words of columns 4*cgroup .. 4*cgroup+3 of row `row`
__device__ __forceinline__ uint4 dropout_words(unsigned long long seed, unsigned int site, unsigned long long row,
                                               unsigned int cgroup) {
  return philox4x32_7(make_uint4((uint32_t)row, (uint32_t)(row >> 32), cgroup, site),
                      make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
}

__device__ __forceinline__ uint32_t word_of(const uint4& w, int i) {
  return i == 0 ? w.x : i == 1 ? w.y : i == 2 ? w.z : w.w;
}

}  // namespace itn
