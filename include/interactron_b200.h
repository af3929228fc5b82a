/*
 * interactron_b200 — C ABI of the B200 (sm_100a) kernels behind the Interactron
 * test-time-adaptation hot path.
 *
 * The reference (allenai/interactron) has no FFI layer: every arithmetic step of
 * the path is an ATen call made from Python.  Each entry point below therefore
 * cites the reference *call site* (file:line under /root/reference) whose
 * arithmetic it replaces.  Conventions:
 *   - plain `extern "C"`, raw device pointers + explicit sizes/strides (in ELEMENTS
 *     unless stated), the CUDA stream as an opaque `void*` (cudaStream_t);
 *   - no allocation, no ownership transfer, no host sync inside any call;
 *   - return 0 on success, negative on error; the message is in itn_last_error();
 *   - re-entrant per stream; every launch is CUDA-graph capturable.
 * Everything is fp32 in memory; GEMMs compute in TF32 on tcgen05 tensor cores.
 */
#ifndef INTERACTRON_B200_H
#define INTERACTRON_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define ITN_OK 0
#define ITN_ERR_ARG -1
#define ITN_ERR_CUDA -2
#define ITN_ERR_UNSUPPORTED -3

/* Last error message of the calling thread ("" if none). */
const char* itn_last_error(void);
/* Library/ABI version and build arch string, e.g. "interactron_b200 0.1 sm_100a". */
const char* itn_version(void);
/* Number of kernels launched by this library since load (all threads). */
long long itn_launch_count(void);

/* ------------------------------------------------------------------ GEMM --- */
/* One matrix operand of a (batched) GEMM.  `major` says which logical dim is
 * contiguous in memory: 0 = the contraction dim K (row-major [rows,K], e.g. an
 * nn.Linear weight [N,K] or an activation [M,K]); 1 = the M (for A) / N (for B)
 * dim (i.e. the matrix is stored [K,rows]).  `ld` is the stride between
 * consecutive indices of the NON-contiguous dim.  sb0/sb1 are the strides of the
 * outer/inner batch index (0 broadcasts the operand over that batch dim). */
typedef struct {
  const float* ptr;
  int major;
  long long ld;
  long long sb0, sb1;
} itn_operand_t;

enum { ITN_ACT_NONE = 0, ITN_ACT_RELU = 1, ITN_ACT_GELU = 2 };
/* ITN_PREC_TF32X3 (default, 0): error-compensated three-pass TF32 (A*B + A_lo*B + A*B_lo with the
 * residual tiles produced in shared memory), ~fp32 accuracy; needed for the 1e-3 parity bar
 * because the inner-loop gradient amplifies single-pass TF32 error to 1-10 %.
 * ITN_PREC_TF32 (1): one tensor-core pass (operands truncated to 10 mantissa bits).
 * ITN_PREC_TF32X3_SPLIT (2): tf32x3 with the two residual products accumulated in tensor-memory columns of
 * their own and added in the epilogue: the main accumulator takes K/8 round-toward-zero accumulates instead
 * of 3K/8 (~fp32 GEMM error; tiles at most 128 wide, slower).  The meta-training step uses it. */
enum { ITN_PREC_TF32X3 = 0, ITN_PREC_TF32 = 1, ITN_PREC_TF32X3_SPLIT = 2 };
enum { ITN_EPI_NONE = 0, ITN_EPI_RELU_MASK = 1, ITN_EPI_GELU_GRAD = 2 };
/* act_pos: 0 = activation right after the bias (default), 1 = after residual/accumulate
 * (ResNet bottleneck: relu(conv(x) + identity)). */

/* C[b0,b1] = epilogue( alpha * A[b0,b1] (MxK) * B[b0,b1]^T (NxK)^T ), batch = nb0*nb1.
 * Epilogue order per element:  v = alpha*acc; v += bias[n]; if (C2) C2 = v;
 * v = act(v) [act_pos 0]; epi (RELU_MASK: v = aux>0 ? v : 0; GELU_GRAD: v *= gelu'(aux));
 * v += residual; if (accumulate) v += C; v = act(v) [act_pos 1]; if (round_out) v = rn_tf32(v); C = v.
 * tcgen05 kind::tf32 TRUNCATES its fp32 operands to 10 mantissa bits, which biases
 * every product by about -4e-4 per operand.  The path therefore keeps every tensor
 * that feeds a GEMM already rounded-to-nearest to TF32 ("TF32-clean"): producers
 * round on store (round_out here, y_r/dx_r in LayerNorm, the softmax kernels,
 * itn_round_tf32 for weights), so the tensor core sees exactly the stored value
 * and the remaining error is unbiased.
 * Replaces: every nn.Linear / F.linear, the 1x1 input_proj conv, bmm/baddbmm of
 * nn.MultiheadAttention and the `q @ k.T`, `att @ v` of CausalSelfAttention on
 * the path, forward and backward (models/detr_models/transformer.py:148-161,
 * 211-232; models/detr_models/detr.py:68-72,299-311; models/gpt.py:43-56,68-78,
 * 197-198; models/transformer.py:49-63; models/new_transformer.py:36-56), and
 * the matching autograd mm/bmm backward nodes of models/interactron.py:51-52. */
typedef struct {
  int M, N, K;
  int nb0, nb1;
  itn_operand_t A, B;
  float* C;              long long ldc,   c_sb0,    c_sb1;
  const float* bias;     long long        bias_sb0, bias_sb1;
  const float* residual; long long ldr,   r_sb0,    r_sb1;
  const float* aux;      long long ldaux, aux_sb0,  aux_sb1;
  float* C2;             long long ldc2,  c2_sb0,   c2_sb1;
  float alpha;
  int act;
  int epi;
  int accumulate;
  int round_out;
  int precision;
  int act_pos;
  int c_pad;        /* 1: columns [N, round_up(N,4)) of every C row belong to C and may be overwritten
                       (with zeros): lets N % 4 != 0 outputs such as attention scores use 128-bit stores.
                       Honoured only for bias-free plain epilogues; otherwise ignored. */
  /* Implicit-GEMM convolution (conv_kh > 0): A.ptr is a channels-last activation [conv_n, conv_h, conv_w, conv_c]
   * and the A tile of k-block (ky, kx, c0..c0+31) is gathered by the TMA unit itself (im2col tensor map: pixel
   * traversal with stride, filter offset ky*dil / kx*dil, zero fill in the padding) - no im2col matrix in HBM.
   * M = conv_n*Ho*Wo, K = conv_kh*conv_kw*conv_c with the (ky, kx, c) column order of itn_im2col_nhwc, conv_c a
   * multiple of 32.  A.major / A.ld / A.sb* are ignored; nb0 = nb1 = 1.  Replaces itn_im2col_nhwc + GEMM for the
   * 3x3 and strided 1x1 convolutions of the frozen trunk (models/detr_models/backbone.py:57-92). */
  int conv_kh, conv_kw, conv_stride, conv_pad, conv_dil;
  int conv_n, conv_h, conv_w, conv_c, conv_ho, conv_wo;
  const float* B_lo; /* optional (may be NULL): B - trunc_tf32(B), element for element at the same offsets and
                       strides as B.ptr (itn_tf32_residual), for a K-major B in tf32x3 mode.  The residual tile of
                       B then arrives by TMA instead of being recomputed in shared memory for every k-block:
                       what static weights are for (+5-10 % on those GEMMs); results are bit-identical. */
} itn_gemm_desc_t;

/* lo = x - trunc_tf32(x): the part of an fp32 value kind::tf32 drops (see itn_gemm_desc_t::B_lo). */
int itn_tf32_residual(const float* x, float* lo, long long n, void* stream);

/* tcgen05/TMA path.  Requires 16-byte aligned operand bases and ld/sb* multiples
 * of 4 elements; returns ITN_ERR_UNSUPPORTED otherwise (use itn_gemm_simt). */
int itn_gemm_tf32(const itn_gemm_desc_t* d, void* stream);
/* 1 if the descriptor satisfies the TMA constraints of itn_gemm_tf32. */
int itn_gemm_tf32_supported(const itn_gemm_desc_t* d);
/* CUDA-core fp32 GEMM with the same descriptor and epilogue, any alignment.
 * Used for the few unaligned / degenerate shapes (N=1 data-grad) on the path. */
int itn_gemm_simt(const itn_gemm_desc_t* d, void* stream);

/* ------------------------------------------------------- fused attention --- */
/* Multi-head attention with the score matrix kept on chip (tcgen05 MMAs with tensor-memory
 * operands, TMA-streamed K/V blocks, online softmax), tf32x3 arithmetic like itn_gemm_tf32.
 * Replaces, forward and backward, the q@k^T -> masked softmax -> @v chains of
 *   models/gpt.py:43-53 (CausalSelfAttention with its all-ones mask = full attention, hd 64),
 *   models/detr_models/transformer.py:154-155 (encoder self-attention, hd 32),
 *   models/detr_models/transformer.py:219-226 (decoder self- and cross-attention; also the
 *   fusion-B layers of models/new_transformer.py:23-25, hd 64),
 * i.e. F.multi_head_attention_forward after the in-projections and before out_proj, with
 * key_padding_mask, eval-mode (no dropout), and the autograd nodes of those ops under
 * models/interactron.py:51-52.  Every tensor is a [B, L, nh*hd] view: element (b, i, h, d) lives at
 * ptr[b*sb + i*ld + h*hd + d] (heads are column blocks of the projection output; q and k may be
 * column halves of one [B, L, 2*nh*hd] buffer).  hd is 32 or 64; pointers 16-byte aligned, ld and
 * sb multiples of 4 elements (otherwise ITN_ERR_UNSUPPORTED; itn_attention_supported tests it).
 *   forward : o = softmax(scale * q k^T + mask) v ;  lse[b,h,i] = log2(sum_j 2^(scale*log2(e)*q_i.k_j))
 *             (log-sum-exp in base 2, what the backward needs to recompute the probabilities).
 *             key_mask (may be NULL): uint8 [B, Lk], 1 = padded key (probability 0).
 *   backward: dq, dk, dv from q, k, v, o, d_o, lse; delta [B, nh, Lq] is scratch (rowsum(d_o * o)).
 *             With dropout the SAME (p, seed, site) as the forward must be given.
 * No atomics: bit-reproducible.  Two launches for the backward (dQ; dK and dV). */
typedef struct {
  int B, nh, hd, Lq, Lk;
  float scale;
  const float* q;   long long q_ld, q_sb;
  const float* k;   long long k_ld, k_sb;
  const float* v;   long long v_ld, v_sb;
  const unsigned char* key_mask;
  float* o;         long long o_ld, o_sb;     /* forward: output; backward: input */
  float* lse;                                  /* [B, nh, Lq]; forward: output; backward: input */
  const float* d_o; long long do_ld, do_sb;   /* backward only from here */
  float* dq;        long long dq_ld, dq_sb;
  float* dk;        long long dk_ld, dk_sb;
  float* dv;        long long dv_ld, dv_sb;
  float* delta;
  /* train()-mode dropout of the attention probabilities (nn.MultiheadAttention(dropout=0.1), gpt.py:51): element
   * (b, h, i, j) is kept iff the Philox word of (seed, site, row (b*nh+h)*Lq+i, column j) >= p*2^32 and scaled
   * by 1/(1-p); see itn_dropout.  drop_p == 0: off.  drop_seed is a DEVICE pointer (graph-replay safe). */
  float drop_p;
  const unsigned long long* drop_seed;
  unsigned int drop_site;
} itn_attention_desc_t;
int itn_attention_supported(const itn_attention_desc_t* d);
int itn_attention_fwd(const itn_attention_desc_t* d, void* stream);
int itn_attention_bwd(const itn_attention_desc_t* d, void* stream);

/* ------------------------------------------------------------- row-wise --- */
/* y = LayerNorm(x) * gamma + beta over the last dim `cols` (eps as given;
 * the reference uses nn.LayerNorm default 1e-5: detr_models/transformer.py:139-140,
 * 198-200; models/gpt.py:64-65,97).  Rows are split in `groups` equal groups;
 * group g uses gamma/beta + g*gb_stride (per-episode fast weights).  mean/rstd
 * ([rows], may be NULL) are saved for the backward.  y_r (may be NULL) receives a
 * TF32-rounded copy of y for GEMM consumers; y itself stays full fp32. */
int itn_layernorm_fwd(const float* x, const float* gamma, const float* beta,
                      float* y, float* y_r, float* mean, float* rstd,
                      long long rows, int cols, int groups, long long gb_stride,
                      float eps, void* stream);
/* LayerNorm forward that also writes y_plus = y + plus, the input of the next attention's q/k projection
 * (`src + pos`, `tgt + query_pos`: detr_models/transformer.py:144-150,207-219), saving the separate add
 * launch and a re-read of y.  plus follows itn_add's broadcast rule: a block of plus_elems values (whole
 * rows) repeated inside each group of plus_group consecutive elements, one block per group
 * (plus_group_stride elements apart; 0 = one block for everything). */
int itn_layernorm_fwd_plus(const float* x, const float* gamma, const float* beta,
                           float* y, float* mean, float* rstd,
                           long long rows, int cols, int groups, long long gb_stride, float eps,
                           const float* plus, float* y_plus, long long plus_elems,
                           long long plus_group, long long plus_group_stride, void* stream);
/* dx = d LayerNorm; dgamma/dbeta (may be NULL; group g at + g*dgb_stride, e.g. a slice
 * of the flat per-episode gradient buffer) are OVERWRITTEN with the per-group sums;
 * dx_r (may be NULL) is a TF32-rounded copy of dx.
 * Replaces native_layer_norm_backward under models/interactron.py:51-52. */
int itn_layernorm_bwd(const float* dy, const float* x, const float* mean,
                      const float* rstd, const float* gamma,
                      float* dx, float* dx_r, float* dgamma, float* dbeta,
                      long long rows, int cols, int groups, long long gb_stride,
                      long long dgb_stride, void* stream);

/* The same backward in ONE launch: dx, dgamma, dbeta and (optionally) dxsum = per-group column sums of dx,
 * i.e. the bias gradient of the linear layer feeding the residual this LayerNorm normalises (post-norm
 * blocks, detr_models/transformer.py:148-160,205-228), which otherwise is a separate column-sum launch.
 * Any of dgamma/dbeta/dxsum may be NULL.  Column sums are combined in a fixed order (deterministic).
 * workspace: itn_layernorm_bwd_fused_workspace(rows, cols, groups) bytes, 16-byte aligned, zero-filled
 * ONCE by the caller (the kernel leaves its counters zero); cols in {128, 256, 512}. */
long long itn_layernorm_bwd_fused_workspace(long long rows, int cols, int groups);
int itn_layernorm_bwd_fused(const float* dy, const float* x, const float* mean,
                            const float* rstd, const float* gamma, float* dx,
                            float* dgamma, float* dbeta, float* dxsum,
                            long long rows, int cols, int groups, long long gb_stride,
                            long long dgb_stride, long long dxsum_stride,
                            void* workspace, long long workspace_bytes, void* stream);

/* In-place row softmax of scale*s + mask over `cols` (row stride ld).
 * key_mask (may be NULL) is uint8 [mask_batches, cols], 1 = padded key (-inf);
 * row r uses mask row (r / rows_per_mask).  round_out: store TF32-rounded
 * probabilities (they only feed GEMMs).  Replaces F.softmax in
 * nn.MultiheadAttention (transformer.py:154,219-226) and models/gpt.py:48-50. */
int itn_softmax_fwd(float* s, long long rows, int cols, long long ld, float scale,
                    const unsigned char* key_mask, long long rows_per_mask,
                    int round_out, void* stream);
/* ds = scale * p * (dp - sum(p*dp)) written in place over dp. */
int itn_softmax_bwd(const float* p, float* dp, long long rows, int cols,
                    long long ld, float scale, int round_out, void* stream);

/* out[g*out_stride + c] = sum_r x[g, r, c]  (bias gradients): x is [groups, rows, cols]
 * with row stride ld and group stride rows*ld; out OVERWRITTEN. */
int itn_colsum(const float* x, float* out, int groups, long long rows, int cols,
               long long ld, long long out_stride, void* stream);

/* --------------------------------------------------------- element-wise --- */
/* out[i] = a[i] + b[(i / a_group) * b_group_stride + i % b_elems]: b is a block of b_elems
 * values broadcast with period b_elems inside each group of a_group consecutive elements of a,
 * with one b block per group (b_group_stride 0 = one block for everything).  Used for
 * `x + pos` and `tgt + query_pos` (reference detr_models/transformer.py:144-150,207-219);
 * optional TF32 rounding of out. */
int itn_add(const float* a, const float* b, float* out, long long n,
            long long b_elems, long long a_group, long long b_group_stride,
            int round_out, void* stream);
/* Strided 2-D copy: dst[r*ldd + c] = src[r*lds + c]; optional TF32 rounding. */
int itn_copy2d(const float* src, long long lds, float* dst, long long ldd,
               long long rows, int cols, int round_out, void* stream);
/* Batched 2-D transpose: dst[g][c][r] = src[g][r][c] (src [groups, rows, cols], group strides in
 * elements).  Builds the W^T twins of the weights so that data-gradient GEMMs (dy W) read both
 * operands K-major: MN-major fp32 operands run ~4x slower through the tensor core. */
int itn_transpose(const float* src, float* dst, int groups, int rows, int cols,
                  long long src_group_stride, long long dst_group_stride, void* stream);
/* dst = rn_tf32(src) (dst may equal src): makes weights / inputs TF32-clean. */
int itn_round_tf32(const float* src, float* dst, long long n, void* stream);
/* y = sigmoid(x)   (detr.py:72 `.sigmoid()`). */
int itn_sigmoid_fwd(const float* x, float* y, long long n, void* stream);
/* dx = dy * y * (1-y). */
int itn_sigmoid_bwd(const float* dy, const float* y, float* dx, long long n,
                    void* stream);
/* Learned loss: per group g of n values, loss[g] = ||x_g||_2 and
 * dx_g = x_g / ||x_g||  (torch.norm + its backward seed, models/interactron.py:50). */
int itn_l2norm_fwd_bwd(const float* x, float* loss, float* dx, int groups, int n,
                       void* stream);

/* train()-mode dropout (the reference trainers call model.train(): engine/interactron_trainer.py:73; nn.Dropout
 * p=0.1 at models/detr_models/transformer.py:156-159,221-230 and models/gpt.py:56,72,195):
 *   out[r, c] = (residual ? residual[r, c] : 0) + x[r, c] * keep(r, c) / (1 - p)
 * keep(r, c) is a pure function of (seed, site, row0 + r, c): Philox4x32 (7 rounds) with counter
 * {row_lo, row_hi, c / 4, site} and key = *seed yields the keep-words of 4 consecutive columns, element kept iff
 * word[c % 4] >= p * 2^32 (csrc/itn_philox.cuh; oracle/philox.py restates it).  The backward pass, the
 * dual-number pass and the fused attention kernels regenerate the mask from the same (seed, site) instead of
 * storing it.  seed is a DEVICE pointer so that captured CUDA graphs draw fresh masks on every replay.
 * Row strides in elements; in place (out == x) allowed. */
int itn_dropout(const float* x, long long ldx, const float* residual, long long ldr, float* out, long long ldo,
                long long rows, int cols, long long row0, float p, const unsigned long long* seed,
                unsigned int site, void* stream);

/* Fused fast-weight step (utils/meta_utils.py:135-142 sgd_step):
 *   theta_out = theta - clip(lr * g, -clip, +clip)
 * theta may be broadcast over `groups` episodes (theta_stride 0); g and
 * theta_out are [groups, n].  theta_out_r (may be NULL) receives the TF32-rounded
 * copy used as GEMM weights.  clip_mask (may be NULL, uint8 [groups,n]) receives
 * 1 where the step was inside the clip band (d step/d g = lr), else 0. */
int itn_sgd_clip_update(const float* theta, long long theta_stride, const float* g,
                        float* theta_out, float* theta_out_r, unsigned char* clip_mask, int groups,
                        long long n, float lr, float clip, void* stream);

/* DETR sine position embedding (detr_models/position_encoding.py:28-48,
 * normalize=True, scale 2*pi, temperature 10000, num_pos_feats=feats per axis).
 * mask uint8 [frames,h,w] (1 = padded) -> pos [frames, h*w, 2*feats] token-major. */
int itn_pos_embed_sine(const unsigned char* mask, float* pos, int frames, int h,
                       int w, int feats, void* stream);

/* ------------------------------------------------------------ frozen trunk --- */
/* Channels-last im2col: dst[(n,ho,wo)][(ky*kw+kx)*C + c] = src[n][ho*stride-pad+ky*dil][wo*stride-pad+kx*dil][c]
 * (zero outside the image); dst row stride ld >= kh*kw*C, extra columns are zeroed.  With the BN-folded
 * weights reshaped to [Cout, kh*kw*Cin] every convolution of the frozen ResNet-50 trunk
 * (reference models/detr_models/backbone.py:57-92, torchvision resnet50) becomes one
 * itn_gemm_tf32 with bias / ReLU / residual fused (k=1,stride=2 is the strided 1x1 gather). */
int itn_im2col_nhwc(const float* src, float* dst, int N, int H, int W, int C, int kh, int kw,
                    int stride, int pad, int dil, int Ho, int Wo, long long ld, void* stream);
/* 3x3 stride-2 pad-1 max pooling, channels-last (C multiple of 4). */
int itn_maxpool3x3s2_nhwc(const float* src, float* dst, int N, int H, int W, int C, int Ho, int Wo,
                          void* stream);

/* ------------------------------------------------------ matcher/criterion --- */
/* HungarianMatcher cost matrix (detr_models/matcher.py:53-71, util/box_ops.py:8-58):
 * for frame f, query q, target t (targets of all frames concatenated, frame f owns
 * [tgt_off[f], tgt_off[f+1])):  C = w_bbox*L1(box_q, box_t) - w_class*softmax(logits_q)[label_t]
 *   - w_giou*GIoU(xyxy(box_q), xyxy(box_t)).  Only the block-diagonal the
 * reference uses is produced: cost is [sum_f Q*T_f] packed frame after frame,
 * row-major [Q, T_f] per frame.  logits [frames,Q,classes], boxes [frames,Q,4]
 * cxcywh, tgt_boxes [T,4] cxcywh, tgt_labels int64 [T], tgt_off int32 [frames+1]. */
int itn_matcher_cost(const float* logits, const float* boxes, const float* tgt_boxes,
                     const long long* tgt_labels, const int* tgt_off, float* cost,
                     int frames, int queries, int classes, float w_class,
                     float w_bbox, float w_giou, void* stream);

/* SetCriterion.forward (detr_models/detr.py:220-265) given the Hungarian assignment, for `groups`
 * independent criterion calls of frames_per_group frames each (one group = one episode's frames, as
 * models/interactron.py:117,131 call it).  logits [groups*frames,Q,classes], boxes [..,Q,4] cxcywh;
 * targets concatenated over all frames (frame f owns [tgt_off[f], tgt_off[f+1])); the matches are
 * (match_row = global row index frame*Q + query, match_tgt = global target index) pairs, group g
 * owning [match_off[g], match_off[g+1]).  Per group:
 *   losses[g] = { loss_ce   weighted CE, weights 1 and background_c on the last class (:111-126),
 *                 class_error  100 - top-1 accuracy on the matched rows (:130-131; 100 if none),
 *                 cardinality_error  mean_f |#(argmax != no-object) - T_f| (:134-146),
 *                 loss_bbox  sum L1 / num_boxes,  loss_giou  sum (1 - GIoU) / num_boxes (:148-167) },
 * num_boxes = max(targets in the group, 1) (:238-242).  If dlogits / dboxes are non-null they receive
 * the gradient of w_ce*loss_ce + w_bbox*loss_bbox + w_giou*loss_giou (what .backward() feeds the
 * detector at models/interactron.py:121-123,133).  scratch: itn_criterion_scratch_bytes() bytes,
 * 16-byte aligned.  Deterministic (fixed-order reductions). */
long long itn_criterion_scratch_bytes(int rows, int n_match, int groups);
int itn_criterion(const float* logits, const float* boxes, const float* tgt_boxes,
                  const long long* tgt_labels, const int* tgt_off, const int* match_row,
                  const int* match_tgt, const int* match_off, int n_match, int groups,
                  int frames_per_group, int queries, int classes, float background_c, float w_ce,
                  float w_bbox, float w_giou, float* losses, float* dlogits, float* dboxes,
                  void* scratch, void* stream);

/* ----------------------------------------- meta-training step (second order) --- */
/* Tangent (forward-mode) companions of the kernels above.  The reference differentiates the inner
 * gradient a second time (`create_graph=True` at models/interactron.py:98-99, then
 * `supervisor_loss.backward()` at :123); here the same numbers come from ONE dual-number re-run of the
 * inner forward+backward along v = -lr * clip_mask * dL_sup/dtheta' (interactron_b200/dual.py), which
 * needs, besides GEMMs, the tangent rules below.  A null "*_dot" input means a zero tangent. */
/* y_dot of LayerNorm: gamma [groups, cols] (row stride g_stride), gamma_dot / beta_dot
 * [groups_dot, cols] (row stride d_stride; per-episode tangents of a shared affine). */
int itn_layernorm_fwd_jvp(const float* x, const float* x_dot, const float* mean, const float* rstd,
                          const float* gamma, const float* gamma_dot, const float* beta_dot,
                          float* y_dot, long long rows, int cols, int groups, long long g_stride,
                          int groups_dot, long long d_stride, void* stream);
/* dx_dot of the LayerNorm backward; gterm (may be NULL, [rows, cols]) receives
 * dy_dot*xhat + dy*xhat_dot whose column sums are the tangent of dgamma. */
int itn_layernorm_bwd_jvp(const float* dy, const float* dy_dot, const float* x, const float* x_dot,
                          const float* mean, const float* rstd, const float* gamma,
                          const float* gamma_dot, float* dx_dot, float* gterm, long long rows, int cols,
                          int groups, long long g_stride, int groups_dot, long long d_stride,
                          void* stream);
/* Softmax backward on dual numbers, in place: dp <- dS, dp_dot <- dS_dot (rows 16-byte aligned and
 * padded to a multiple of 4 columns, as the attention score buffers are). */
int itn_softmax_bwd_jvp(const float* p, const float* p_dot, float* dp, float* dp_dot, long long rows,
                        int cols, long long ld, float scale, void* stream);
/* y *= (ref > 0): tangent of ReLU / of the ReLU-mask backward epilogue. */
int itn_mask_mul(float* y, const float* ref, long long n, void* stream);
/* out = scale * x * mask (mask uint8 from itn_sgd_clip_update): the derivative of the clipped SGD step
 * wrt g applied to dL/dtheta' (utils/meta_utils.py:141: lr inside the clip band, 0 outside). */
int itn_mul_mask_u8(const float* x, const unsigned char* mask, float scale, float* out, long long n,
                    void* stream);
/* y = raw * gelu'(aux); y_dot = raw_dot * gelu'(aux) + raw * gelu''(aux) * aux_dot (exact-erf GELU,
 * models/gpt.py:70).  y_dot may be NULL. */
int itn_gelu_grad_dual(const float* raw, const float* raw_dot, const float* aux, const float* aux_dot,
                       float* y, float* y_dot, long long n, void* stream);
/* out = dy_dot*y*(1-y) + dy*(1-2y)*y_dot: tangent of itn_sigmoid_bwd. */
int itn_sigmoid_bwd_jvp(const float* dy, const float* dy_dot, const float* y, const float* y_dot,
                        float* out, long long n, void* stream);
/* Tangent of itn_l2norm_fwd_bwd: n_dot[g] = <d_g, x_dot_g>, d_dot_g = (x_dot_g - d_g*n_dot[g]) / nrm[g]. */
int itn_l2norm_jvp(const float* x_dot, const float* nrm, const float* d, float* n_dot, float* d_dot,
                   int groups, int n, void* stream);

/* ---- trainer step over the flat buffers -------------------------------------
 * Replaces, for the meta-training loop of engine/interactron_trainer.py:106-110,
 *   torch.nn.utils.clip_grad_norm_(model.parameters(), GRAD_NORM_CLIP)
 *   detector_optimizer.step(); supervisor_optimizer.step()     (torch.optim.Adam, default betas/eps)
 *   detector_optimizer.zero_grad(); supervisor_optimizer.zero_grad()
 * itn_sumsq_partials: partial[i] = sum of squares of a fixed slice of g[0..n); *n_partials (host int)
 * receives how many were written (<= max_partials, <= 1184).  Fixed summation order.
 * itn_clip_adam_step: one segment (w, g, m, v of n elements; g may point inside the buffer the
 * partials were taken over).  total = sqrt(sum(partial)) is the global gradient norm (written to
 * norm_out if not NULL); g is scaled by min(1, max_norm / (total + 1e-6)) when max_norm > 0 and
 * partial != NULL; then m, v, w are updated exactly as torch.optim.Adam does at step `step` (1-based)
 * with learning rate lr (hyper-parameters are doubles, like torch's Python scalars: 1 - beta2 must not be
 * formed in fp32); zero_grad != 0 clears g afterwards.
 * An element of g whose bit pattern is ITN_NO_GRAD_BITS (a quiet NaN with a payload arithmetic never
 * produces) means "no gradient" (torch: .grad is None): it adds nothing to the norm and its w, m, v
 * are left untouched, as clip_grad_norm_ / Adam skip such parameters.  Genuine NaNs propagate. */
#define ITN_NO_GRAD_BITS 0x7FC0DEADu
int itn_sumsq_partials(const float* g, long long n, float* partial, int max_partials, int* n_partials,
                       void* stream);
int itn_clip_adam_step(float* w, float* g, float* m, float* v, long long n, const float* partial,
                       int n_partials, float max_norm, double lr, double beta1, double beta2, double eps,
                       int step, int zero_grad, float* norm_out, void* stream);

/* Windowed checkpoint average (engine/interactron_trainer.py:48-57: `saved[k] = w * v` on the first record,
 * `saved[k] += w * v` afterwards, w = 1/SAVE_WINDOW over the last SAVE_WINDOW epochs) on a flat weight buffer:
 * acc = (first ? 0 : acc) + w * x, element-wise in torch's operation order (product rounded, then the sum:
 * bit-identical to the reference's two torch ops).  One launch per flat buffer per recorded epoch. */
int itn_ckpt_accumulate(const float* x, float* acc, long long n, float w, int first, void* stream);

/* ---- evaluator post-processing ------------------------------------------------
 * Replaces, per image, engine/random_policy_evaluator.py:65-78 (= engine/interactive_evaluator.py:71-84):
 *   pred_scores, pred_cats = logits.softmax(-1).max(-1); drop pred_cats == background_class;
 *   boxes cxcywh -> xyxy (detr_models/util/box_ops.py:8-12); torchvision.ops.nms(boxes, scores, iou_threshold).
 * logits [images, queries, classes], boxes [images, queries, 4] (cxcywh) ->
 *   count [images]: detections kept; for k < count[i], in decreasing score order (ties: lower query first):
 *   keep_idx [images, queries] (query index, -1 beyond count), score, cat (int32), xyxy [images, queries, 4].
 * queries <= 128.  The keep/suppress decisions are bit-identical to torchvision's CPU nms on the same
 * xyxy boxes and scores (same operation order, no fused multiply-add). */
int itn_detect_postprocess(const float* logits, const float* boxes, int images, int queries, int classes,
                           int background_class, float iou_threshold, int* count, int* keep_idx,
                           float* score, int* cat, float* xyxy, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* INTERACTRON_B200_H */
