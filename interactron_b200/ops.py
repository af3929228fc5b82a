"""Tensor-level wrappers over the C ABI (`CudaOps`).

Every method takes/returns torch CUDA tensors (fp32) and launches exactly the
hand-written sm_100a kernels of libinteractron_b200.so on torch's current
stream; torch is used only for device memory and views.  The orchestration
modules (`detr_t`, `fusion_a`, `fusion_b`, `episode`) are written against this
small interface; `oracle/sim_ops.py` implements the same interface with plain
torch CPU ops and exists only so the tests can check the hand-derived backward
passes against the reference's autograd without a GPU.
"""
import ctypes as C
import os

import torch

from . import _lib

ACT = {None: 0, "none": 0, "relu": 1, "gelu": 2}
PRECISION = {"tf32x3": 0, "tf32": 1}
EPI = {None: 0, "none": 0, "relu_mask": 1, "gelu_grad": 2}


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _as4d(t):
    while t.dim() < 4:
        t = t.unsqueeze(0)
    if t.dim() != 4:
        raise ValueError(f"at most two batch dims supported, got shape {tuple(t.shape)}")
    return t


class CudaOps:
    """The product backend.  Requires a CUDA device and the built shared library."""

    name = "cuda"

    def __init__(self, device="cuda:0"):
        if not torch.cuda.is_available():
            raise _lib.ItnError("interactron_b200 needs a CUDA device (sm_100a); there is no CPU path")
        self.lib = _lib.load()
        self.device = torch.device(device)
        self.force_simt = os.environ.get("ITN_FORCE_SIMT", "0") == "1"
        # "tf32x3": error-compensated 3-pass TF32 (~fp32 accuracy, the parity mode, default);
        # "tf32": single pass with TF32-clean (round-to-nearest at the producer) operands.
        self.precision = os.environ.get("ITN_GEMM_PRECISION", "tf32x3")
        # tf32x3 with the residual products in their own TMEM accumulator (ITN_PREC_TF32X3_SPLIT): ~fp32 GEMM
        # error (measured 3.9e-7 / 5.8e-7 at K = 256 / 2048 against torch fp32's 2.9e-7 / 8.1e-7), tiles at most
        # 128 wide so 8-20 % slower.  Off by default; flip per instance or with ITN_GEMM_SPLITACC=1.
        self.split_acc = os.environ.get("ITN_GEMM_SPLITACC", "0") != "0"
        if self.precision not in PRECISION:
            raise ValueError(f"ITN_GEMM_PRECISION must be one of {sorted(PRECISION)}")
        self.n_tf32 = 0
        self.n_simt = 0
        self.n_split_k = 0
        self.split_k = os.environ.get("ITN_SPLIT_K", "1") != "0"
        # static GEMM weights registered with their tf32 residuals (register_presplit): [(base, end, lo, src)]
        self._presplit = []
        self.presplit = os.environ.get("ITN_PRESPLIT", "1") != "0"
        self.n_presplit = 0
        # L2 budget (MiB) for the score tensors of one attention chunk (layers._l2_chunks); 0 = one pass
        self.attn_l2_mb = int(os.environ.get("ITN_ATTN_L2_MB", "0"))
        # fused tcgen05 attention (itn_attention_fwd/bwd); ITN_FUSED_ATTN=0 keeps the unfused
        # QK^T -> softmax -> PV chain through HBM (cross-check / dual-number path)
        self.fused_attention = os.environ.get("ITN_FUSED_ATTN", "1") != "0"
        # convolutions as implicit GEMMs (TMA im2col tensor maps); ITN_IMPLICIT_CONV=0: explicit itn_im2col_nhwc + GEMM
        self.implicit_conv = os.environ.get("ITN_IMPLICIT_CONV", "1") != "0"
        # the 4-channel (zero-padded RGB) 7x7 stem as an implicit GEMM too (one TMA im2col load per filter tap);
        # ITN_IMPLICIT_STEM=0: explicit im2col of the stem (5.5 GB written and re-read at 62 episodes)
        self.implicit_stem = os.environ.get("ITN_IMPLICIT_STEM", "1") != "0"
        self.n_attn = 0
        # LayerNorm backward in one launch (dx + dgamma/dbeta + the bias gradient colsum(dx)); ITN_FUSED_LN_BWD=0:
        # the dx kernel, the dgamma/dbeta kernel and a separate column sum
        self.fused_ln_bwd = os.environ.get("ITN_FUSED_LN_BWD", "1") != "0"
        self._ln_ws = None
        self._ln_ws_keep = []
        # LayerNorm forward that also emits y + pos / y + query_pos for the next attention (ITN_FUSED_LN_PLUS=0: add kernel)
        self.fused_ln_plus = os.environ.get("ITN_FUSED_LN_PLUS", "1") != "0"

    # ------------------------------------------------------------ plumbing
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def empty(self, *shape):
        return torch.empty(*shape, dtype=torch.float32, device=self.device)

    def zeros(self, *shape):
        return torch.zeros(*shape, dtype=torch.float32, device=self.device)

    def launch_count(self):
        return int(self.lib.itn_launch_count())

    @property
    def precision_key(self):
        """What a captured graph depends on (GEMM arithmetic mode)."""
        return self.precision + ("+split" if self.split_acc else "")

    def _prec(self):
        return 2 if (self.split_acc and self.precision == "tf32x3") else PRECISION[self.precision]

    @property
    def _clean(self):
        """True when GEMM operands must be stored TF32-rounded (single-pass mode only)."""
        return self.precision == "tf32"

    # ------------------------------------------------- pre-split weights (itn_gemm_desc_t::B_lo)
    def register_presplit(self, w):
        """Keep `lo = w - trunc_tf32(w)` next to a persistent weight buffer `w` (contiguous).  Every tf32x3
        GEMM whose K-major B operand is a view into `w` then receives the matching view of `lo` as B_lo, so the
        residual tile of the weights arrives by TMA instead of being recomputed per k-block.  Call again after
        `w` changed in place (same tensor: the residual is refreshed in place; bit-identical results either
        way, the kernel evaluates the same expression)."""
        if not self.presplit or self.precision != "tf32x3":
            return
        assert w.is_contiguous() and w.dtype == torch.float32
        base = w.data_ptr()
        for ent in self._presplit:
            if ent[0] == base and ent[3] is w:
                _lib.check(self.lib.itn_tf32_residual(_ptr(w), _ptr(ent[2]), w.numel(), self._stream()))
                return
        lo = torch.empty_like(w)
        _lib.check(self.lib.itn_tf32_residual(_ptr(w), _ptr(lo), w.numel(), self._stream()))
        # the entry holds `w` itself: its address range cannot be handed to another tensor while it is listed
        self._presplit.append((base, base + w.numel() * 4, lo, w))
        self._presplit.sort(key=lambda e: e[0])

    def _b_lo_ptr(self, ptr):
        ents = self._presplit
        lo_i, hi_i = 0, len(ents)
        while lo_i < hi_i:                       # last entry with base <= ptr
            mid = (lo_i + hi_i) // 2
            if ents[mid][0] <= ptr:
                lo_i = mid + 1
            else:
                hi_i = mid
        if lo_i == 0:
            return 0
        base, end, lo, _ = ents[lo_i - 1]
        return lo.data_ptr() + (ptr - base) if ptr < end else 0

    # ---------------------------------------------------------------- GEMM
    @staticmethod
    def _operand(t4, nb0, nb1, is_a):
        """t4: [b0,b1,M,K] for A, [b0,b1,K,N] for B (any strides).  -> (Operand, keepalive)."""
        if is_a:
            rows, k = t4.shape[2], t4.shape[3]
            s_rows, s_k = t4.stride(2), t4.stride(3)
        else:
            k, rows = t4.shape[2], t4.shape[3]
            s_k, s_rows = t4.stride(2), t4.stride(3)
        k_contig = k == 1 or s_k == 1
        r_contig = rows == 1 or s_rows == 1
        pad4 = lambda v: max(4, (v + 3) // 4 * 4)
        ld_k = s_rows if rows > 1 else pad4(k)      # ld if K is the contiguous dim
        ld_r = s_k if k > 1 else pad4(rows)         # ld if rows is the contiguous dim
        if k_contig and ld_k % 4 == 0:
            major, ld = 0, ld_k
        elif r_contig and ld_r % 4 == 0:
            major, ld = 1, ld_r
        elif k_contig:
            major, ld = 0, ld_k                     # unaligned: the SIMT kernel takes it
        elif r_contig:
            major, ld = 1, ld_r
        else:
            t4 = t4.contiguous()
            return CudaOps._operand(t4, nb0, nb1, is_a)
        op = _lib.Operand()
        op.ptr = t4.data_ptr()
        op.major = major
        op.ld = ld
        op.sb0 = t4.stride(0) if t4.shape[0] > 1 else 0
        op.sb1 = t4.stride(1) if t4.shape[1] > 1 else 0
        return op, t4

    def matmul(self, a, b, *, bias=None, act=None, residual=None, out=None, out_pre=None,
               alpha=1.0, accumulate=False, epi=None, aux=None, rnd=False, act_after_residual=False,
               out_pad=False, _nosplit=False):
        """out = epilogue(alpha * a @ b), a [..,M,K], b [..,K,N]; see itn_gemm_desc_t.
        rnd: store `out` rounded to TF32 (set when `out` only feeds further GEMMs).
        out_pad: `out` is a [..., :N] view of rows padded to a multiple of 4 columns and the pad may
        be overwritten with zeros (attention score matrices): enables 128-bit stores for N % 4 != 0."""
        rank = max(a.dim(), b.dim(), 2)
        a4, b4 = _as4d(a), _as4d(b)
        M, K = a4.shape[2], a4.shape[3]
        K2, N = b4.shape[2], b4.shape[3]
        if K != K2:
            raise ValueError(f"matmul inner dims differ: {tuple(a.shape)} @ {tuple(b.shape)}")
        nb0 = max(a4.shape[0], b4.shape[0])
        nb1 = max(a4.shape[1], b4.shape[1])
        plain = (bias is None and act is None and residual is None and out_pre is None and epi is None and aux is None
                 and alpha == 1.0 and not out_pad and not (rnd and self._clean))
        if plain and not _nosplit and out is not None and self.split_k:
            S = self._split_k_factor(a4, b4, out, M, N, K, nb0, nb1)
            if S > 1:
                return self._matmul_split_k(a4, b4, out, S, accumulate)
        for t in (a4, b4):
            if t.shape[0] not in (1, nb0) or t.shape[1] not in (1, nb1):
                raise ValueError(f"batch dims not broadcastable: {tuple(a.shape)} @ {tuple(b.shape)}")
        if out is None:
            out = self.empty(nb0, nb1, M, N)
            ret = out.reshape(out.shape[4 - rank:])
        else:
            ret = out
        d = _lib.GemmDesc()
        d.M, d.N, d.K, d.nb0, d.nb1 = M, N, K, nb0, nb1
        d.A, keep_a = self._operand(a4, nb0, nb1, True)
        d.B, keep_b = self._operand(b4, nb0, nb1, False)
        if self._presplit and d.B.major == 0 and self.precision == "tf32x3" and keep_b is b4:
            blo = self._b_lo_ptr(b4.data_ptr())
            if blo:
                d.B_lo = blo
                self.n_presplit += 1

        def mat(t, what):
            t4 = _as4d(t)
            if tuple(t4.shape[2:]) != (M, N) or (N > 1 and t4.stride(3) != 1):
                raise ValueError(f"{what} must be [..,{M},{N}] with unit inner stride, got "
                                 f"{tuple(t.shape)} strides {t.stride()}")
            sb0 = t4.stride(0) if t4.shape[0] > 1 else 0
            sb1 = t4.stride(1) if t4.shape[1] > 1 else 0
            return t4.data_ptr(), t4.stride(2), sb0, sb1

        d.C, d.ldc, d.c_sb0, d.c_sb1 = mat(out, "out")
        o4 = _as4d(out)
        if o4.shape[0] != nb0 or o4.shape[1] != nb1:
            raise ValueError("out batch dims must match the broadcast batch")
        if bias is not None:
            b3 = bias
            while b3.dim() < 3:
                b3 = b3.unsqueeze(0)
            if b3.shape[-1] != N or (N > 1 and b3.stride(-1) != 1):
                raise ValueError("bias must be [..,N] contiguous in N")
            d.bias = b3.data_ptr()
            d.bias_sb0 = b3.stride(0) if b3.shape[0] > 1 else 0
            d.bias_sb1 = b3.stride(1) if b3.shape[1] > 1 else 0
        if residual is not None:
            d.residual, d.ldr, d.r_sb0, d.r_sb1 = mat(residual, "residual")
        if aux is not None:
            d.aux, d.ldaux, d.aux_sb0, d.aux_sb1 = mat(aux, "aux")
        if out_pre is not None:
            d.C2, d.ldc2, d.c2_sb0, d.c2_sb1 = mat(out_pre, "out_pre")
        d.alpha = float(alpha)
        d.act = ACT[act]
        d.epi = EPI[epi]
        d.accumulate = 1 if accumulate else 0
        d.round_out = 1 if (rnd and self._clean) else 0
        d.precision = self._prec()
        d.act_pos = 1 if act_after_residual else 0
        d.c_pad = 1 if out_pad else 0
        if not self.force_simt and self.lib.itn_gemm_tf32_supported(C.byref(d)):
            _lib.check(self.lib.itn_gemm_tf32(C.byref(d), self._stream()))
            self.n_tf32 += 1
        else:
            _lib.check(self.lib.itn_gemm_simt(C.byref(d), self._stream()))
            self.n_simt += 1
        del keep_a, keep_b
        return ret

    # ------------------------------------------------------------ fused attention
    @staticmethod
    def _heads_view(d, prefix, t, nh, hd):
        """[B, L, nh*hd] view with unit inner stride -> (ptr, row stride, batch stride) fields of the descriptor."""
        if t.dim() != 3 or t.shape[2] != nh * hd or t.stride(2) != 1:
            raise ValueError(f"attention operand must be [B, L, {nh * hd}] with unit inner stride, got "
                             f"{tuple(t.shape)} strides {t.stride()}")
        setattr(d, prefix, t.data_ptr())
        setattr(d, prefix + "_ld" if prefix != "d_o" else "do_ld", t.stride(1))
        setattr(d, prefix + "_sb" if prefix != "d_o" else "do_sb", t.stride(0) if t.shape[0] > 1 else 0)

    def _attention_desc(self, q, k, v, o, lse, nh, scale, kmask, drop=None):
        B, Lq, D = q.shape
        Lk = k.shape[1]
        hd = D // nh
        d = _lib.AttentionDesc()
        d.B, d.nh, d.hd, d.Lq, d.Lk, d.scale = B, nh, hd, Lq, Lk, float(scale)
        for name, t in (("q", q), ("k", k), ("v", v), ("o", o)):
            self._heads_view(d, name, t, nh, hd)
        if k.shape != v.shape or k.shape[0] != B or o.shape != q.shape:
            raise ValueError("attention: shape mismatch between q/k/v/o")
        if kmask is not None:
            assert kmask.dtype == torch.uint8 and kmask.is_contiguous() and tuple(kmask.shape) == (B, Lk)
            d.key_mask = kmask.data_ptr()
        assert lse.is_contiguous() and lse.numel() == B * nh * Lq
        d.lse = lse.data_ptr()
        if drop is not None:                         # (p, device seed tensor, site): train()-mode dropout of the probabilities
            p_, seed, site = drop
            assert seed.is_cuda and seed.dtype == torch.int64 and seed.numel() == 1
            d.drop_p, d.drop_seed, d.drop_site = float(p_), seed.data_ptr(), int(site)
        return d

    def attention_supported(self, q, k, v, nh):
        """True when the fused tcgen05 attention kernels take these views (hd 32/64, 16-byte aligned rows)."""
        if self.force_simt or not self.fused_attention or self._clean:
            return False
        hd = q.shape[-1] // nh
        if hd not in (32, 64):
            return False
        for t in (q, k, v):
            if t.dim() != 3 or t.stride(2) != 1 or t.data_ptr() % 16 or t.stride(1) % 4 or \
                    (t.shape[0] > 1 and t.stride(0) % 4):
                return False
        return True

    def attention_fwd(self, q, k, v, nh, scale, kmask=None, drop=None):
        """o = softmax(scale q k^T + key mask) v per (batch, head), scores kept on chip.
        q [B,Lq,nh*hd], k/v [B,Lk,nh*hd] (strided views ok) -> (o [B,Lq,nh*hd], lse [B,nh,Lq] base-2 log-sum-exp)."""
        B, Lq, D = q.shape
        o = self.empty(B, Lq, D)
        lse = self.empty(B, nh, Lq)
        d = self._attention_desc(q, k, v, o, lse, nh, scale, kmask, drop)
        _lib.check(self.lib.itn_attention_fwd(C.byref(d), self._stream()))
        self.n_attn += 1
        return o, lse

    def attention_bwd(self, dO, q, k, v, o, lse, nh, scale, kmask, dq, dk, dv, drop=None):
        """Writes dq/dk/dv (views like q/k/v) given dO [B,Lq,nh*hd]; scores are recomputed from lse
        (and the dropout mask from `drop`, which must be the forward's)."""
        B, Lq, D = q.shape
        d = self._attention_desc(q, k, v, o, lse, nh, scale, kmask, drop)
        hd = D // nh
        for name, t in (("d_o", dO), ("dq", dq), ("dk", dk), ("dv", dv)):
            self._heads_view(d, name, t, nh, hd)
        delta = self.empty(B, nh, Lq)
        d.delta = delta.data_ptr()
        _lib.check(self.lib.itn_attention_bwd(C.byref(d), self._stream()))
        self.n_attn += 2

    # ------------------------------------------------------------ implicit-GEMM convolution
    def conv_gemm(self, x, w, kh, kw, stride, pad, dil, bias=None, act=None, residual=None, act_after_residual=False):
        """Convolution of channels-last x [N,H,W,C] with w [Cout, kh*kw*C] (columns in (ky, kx, c) order) as ONE
        tcgen05 GEMM whose A tiles the TMA unit gathers in im2col mode (no patch matrix in HBM) -> [N*Ho*Wo, Cout].
        C % 32 == 0, or C == 4 (the zero-padded RGB stem: one im2col load per filter tap, 8 taps per k-block);
        bias / activation / residual fused as in `matmul`."""
        assert x.is_contiguous() and x.dim() == 4 and w.dim() == 2 and w.stride(1) == 1
        N, H, W_, Cc = x.shape
        Ho = (H + 2 * pad - dil * (kh - 1) - 1) // stride + 1
        Wo = (W_ + 2 * pad - dil * (kw - 1) - 1) // stride + 1
        M, K, Nn = N * Ho * Wo, kh * kw * Cc, w.shape[0]
        assert (Cc % 32 == 0 or Cc == 4) and w.shape[1] >= K
        out = self.empty(M, Nn)
        d = _lib.GemmDesc()
        d.M, d.N, d.K, d.nb0, d.nb1 = M, Nn, K, 1, 1
        d.A.ptr, d.A.major, d.A.ld = x.data_ptr(), 0, K
        d.B, keep_b = self._operand(_as4d(w.t()), 1, 1, False)
        if self._presplit and d.B.major == 0 and self.precision == "tf32x3":
            blo = self._b_lo_ptr(w.data_ptr())
            if blo:
                d.B_lo = blo
                self.n_presplit += 1
        d.C, d.ldc = out.data_ptr(), Nn
        if bias is not None:
            assert bias.is_contiguous() and bias.numel() == Nn
            d.bias = bias.data_ptr()
        if residual is not None:
            assert residual.is_contiguous() and residual.numel() == M * Nn
            d.residual, d.ldr = residual.data_ptr(), Nn
        d.alpha, d.act, d.precision = 1.0, ACT[act], self._prec()
        d.act_pos = 1 if act_after_residual else 0
        d.conv_kh, d.conv_kw, d.conv_stride, d.conv_pad, d.conv_dil = kh, kw, stride, pad, dil
        d.conv_n, d.conv_h, d.conv_w, d.conv_c, d.conv_ho, d.conv_wo = N, H, W_, Cc, Ho, Wo
        _lib.check(self.lib.itn_gemm_tf32(C.byref(d), self._stream()))
        self.n_tf32 += 1
        del keep_b
        return out, Ho, Wo

    # split-K: weight-gradient GEMMs at few episodes per step (dW[256,256] = dy^T[256,3610] x[3610,256]) are
    # 4-32 output tiles with a 57-113 k-block chain each: 3-20 % of the 148 SMs busy for 40-130 us.  The K
    # range is cut into S equal chunks that run as one more batch dimension of the same kernel (strided
    # views, nothing is copied), and the S partial products are summed in a fixed order by colsum.
    # Measured (bench.py --workload meta_*, 2 episodes/step): +5-6 %.  K < 2048 (the per-episode gradients of
    # a 1-episode predict(), K = 1805) is left alone: there the second launch costs more than it saves (-4 %).
    def _split_k_factor(self, a4, b4, out, M, N, K, nb0, nb1):
        if nb0 != 1 or K < 2048 or M < 32 or N < 32:
            return 1
        o4 = _as4d(out)
        if o4.shape[0] != 1 or o4.shape[1] != nb1 or (N > 1 and o4.stride(3) != 1) or (M > 1 and o4.stride(2) != N):
            return 1
        tiles = ((M + 127) // 128) * ((N + 127) // 128) * nb1
        if tiles > 37:
            return 1
        # a chunk boundary must keep every operand TMA-legal: 16-byte aligned batch stride when K is the
        # contiguous dim (MN-major operands step by whole rows and are always fine)
        need4 = (a4.stride(3) == 1 and a4.stride(2) != 1) or (b4.stride(2) == 1 and b4.stride(3) != 1)
        best = 1
        for S in range(2, 65):
            if K % S or K // S < 64 or tiles * S > 296 or (need4 and (K // S) % 4):
                continue
            best = S
        return best if best >= 4 else 1

    def _matmul_split_k(self, a4, b4, out, S, accumulate):
        K = a4.shape[3]
        Kc = K // S
        nb1 = max(a4.shape[1], b4.shape[1])
        M, N = a4.shape[2], b4.shape[3]
        a5 = a4[0].unflatten(-1, (S, Kc)).movedim(-2, 1)         # [nb1|1, S, M, Kc]
        b5 = b4[0].unflatten(-2, (S, Kc))                        # [nb1|1, S, Kc, N]
        extra = 1 if accumulate else 0
        part = self.empty(nb1, S + extra, M, N)
        self.matmul(a5, b5, out=part[:, :S], _nosplit=True)
        o3 = _as4d(out)[0]                                       # [nb1, M, N], rows contiguous
        if accumulate:
            part[:, S].copy_(o3)
        self.colsum(part.view(nb1, S + extra, M * N), out=o3.flatten(1))
        self.n_split_k += 1
        return out

    # ------------------------------------------------------------ row-wise
    def layernorm_fwd(self, x, gamma, beta, eps=1e-5):
        """x [rows, D] contiguous; gamma/beta [G, D] (G divides rows).
        -> y (fp32), y_r (TF32-rounded copy for GEMM consumers), mean, rstd."""
        rows, cols = x.shape
        g2, b2 = gamma.reshape(-1, cols), beta.reshape(-1, cols)
        groups = g2.shape[0]
        assert x.is_contiguous() and g2.stride(1) == 1 and b2.stride(1) == 1
        assert groups == 1 or g2.stride(0) == b2.stride(0)
        y = self.empty(rows, cols)
        y_r = self.empty(rows, cols) if self._clean else None
        mean, rstd = self.empty(rows), self.empty(rows)
        _lib.check(self.lib.itn_layernorm_fwd(_ptr(x), _ptr(g2), _ptr(b2), _ptr(y), _ptr(y_r), _ptr(mean),
                                              _ptr(rstd), rows, cols, groups, g2.stride(0) if groups > 1 else 0,
                                              eps, self._stream()))
        return y, (y_r if y_r is not None else y), mean, rstd

    def layernorm_fwd_plus(self, x, gamma, beta, plus, shape, eps=1e-5):
        """layernorm_fwd that also returns y + plus, with `add`'s broadcast rule applied to y viewed as `shape`
        ([G, n, cols]) - one launch instead of LayerNorm + add.  -> y, y_r, mean, rstd, y_plus [rows, cols]."""
        if self._clean or not self.fused_ln_plus:
            y, y_r, mean, rstd = self.layernorm_fwd(x, gamma, beta, eps)
            return y, y_r, mean, rstd, self.add(y.view(shape), plus, rnd=True).view(x.shape)
        rows, cols = x.shape
        g2, b2 = gamma.reshape(-1, cols), beta.reshape(-1, cols)
        groups = g2.shape[0]
        assert x.is_contiguous() and g2.stride(1) == 1 and b2.stride(1) == 1
        assert groups == 1 or g2.stride(0) == b2.stride(0)
        n = x.numel()
        grouped = plus.dim() >= 2 and plus.shape[0] == shape[0] and plus.shape[0] > 1
        if grouped and (n != plus.numel() or not plus.is_contiguous()):
            G = shape[0]
            assert plus[0].is_contiguous()
            b_elems, a_group, b_gs = plus.numel() // G, n // G, plus.stride(0)
        else:
            assert plus.is_contiguous()
            b_elems, a_group, b_gs = plus.numel(), n, 0
        assert n % b_elems == 0
        y, y_plus = self.empty(rows, cols), self.empty(rows, cols)
        mean, rstd = self.empty(rows), self.empty(rows)
        _lib.check(self.lib.itn_layernorm_fwd_plus(_ptr(x), _ptr(g2), _ptr(b2), _ptr(y), _ptr(mean), _ptr(rstd), rows,
                                                   cols, groups, g2.stride(0) if groups > 1 else 0, eps, _ptr(plus),
                                                   _ptr(y_plus), b_elems, a_group, b_gs, self._stream()))
        return y, y, mean, rstd, y_plus

    def layernorm_bwd(self, dy, x, mean, rstd, gamma, dgamma=None, dbeta=None, dxsum=None):
        """-> (dx, dx_r).  dgamma/dbeta: optional [groups, cols] views
        (row stride arbitrary, e.g. slices of the flat gradient buffer) that receive the
        per-group affine gradients.  dxsum (fused_ln_bwd only): optional [groups, cols] view that receives
        the per-group column sums of dx (the bias gradient of the linear layer feeding the residual)."""
        rows, cols = x.shape
        g2 = gamma.reshape(-1, cols)
        groups = g2.shape[0] if dgamma is None else dgamma.shape[0]
        if dgamma is None and dxsum is not None:
            groups = dxsum.shape[0]
        assert dy.is_contiguous() and x.is_contiguous() and g2.stride(1) == 1
        assert g2.shape[0] in (1, groups)
        dx = self.empty(rows, cols)
        dx_r = self.empty(rows, cols) if self._clean else None
        stride = 0
        if dgamma is not None:
            assert dgamma.shape == dbeta.shape == (groups, cols)
            stride = dgamma.stride(0) if groups > 1 else cols
            assert groups == 1 or dbeta.stride(0) == stride
        gb_stride = g2.stride(0) if g2.shape[0] > 1 else 0
        if self.fused_ln_bwd and not self._clean and cols in (128, 256, 512) and (dgamma is not None or dxsum is not None):
            xs_stride = 0
            if dxsum is not None:
                assert dxsum.shape == (groups, cols) and dxsum.stride(1) == 1
                xs_stride = dxsum.stride(0) if groups > 1 else cols
            need = int(self.lib.itn_layernorm_bwd_fused_workspace(rows, cols, groups))
            if self._ln_ws is None or self._ln_ws.numel() < need:
                # allocated (and zeroed) outside of graph capture by the eager warm-up pass; grows monotonically.
                # Captured graphs keep the address of the workspace they were recorded with: outgrown ones stay alive.
                if torch.cuda.is_current_stream_capturing():
                    raise RuntimeError("layernorm_bwd: workspace must be sized by an eager pass before graph capture")
                self._ln_ws = torch.zeros(max(need, 8 << 20), dtype=torch.uint8, device=self.device)
                self._ln_ws_keep.append(self._ln_ws)
            _lib.check(self.lib.itn_layernorm_bwd_fused(_ptr(dy), _ptr(x), _ptr(mean), _ptr(rstd), _ptr(g2), _ptr(dx),
                                                        _ptr(dgamma), _ptr(dbeta), _ptr(dxsum), rows, cols, groups,
                                                        gb_stride, stride, xs_stride, _ptr(self._ln_ws),
                                                        self._ln_ws.numel(), self._stream()))
            return dx, dx
        assert dxsum is None, "dxsum needs the fused LayerNorm backward"
        _lib.check(self.lib.itn_layernorm_bwd(_ptr(dy), _ptr(x), _ptr(mean), _ptr(rstd), _ptr(g2), _ptr(dx),
                                              _ptr(dx_r), _ptr(dgamma), _ptr(dbeta), rows, cols, groups,
                                              gb_stride, stride, self._stream()))
        return dx, (dx_r if dx_r is not None else dx)

    def softmax_(self, s, cols, scale, key_mask=None, rows_per_mask=1):
        """In-place softmax over the first `cols` entries of the last dim of contiguous `s`
        (stored TF32-rounded: probabilities only feed GEMMs)."""
        assert s.is_contiguous()
        ld = s.shape[-1]
        rows = s.numel() // ld
        if key_mask is not None:
            assert key_mask.dtype == torch.uint8 and key_mask.is_contiguous() and key_mask.shape[-1] == cols
        _lib.check(self.lib.itn_softmax_fwd(_ptr(s), rows, cols, ld, float(scale), _ptr(key_mask),
                                            rows_per_mask, 1 if self._clean else 0, self._stream()))
        return s

    def softmax_bwd_(self, p, dp, cols, scale):
        """dp <- scale * p * (dp - rowsum(p*dp)) in place."""
        assert p.is_contiguous() and dp.is_contiguous() and p.shape == dp.shape
        ld = p.shape[-1]
        rows = p.numel() // ld
        _lib.check(self.lib.itn_softmax_bwd(_ptr(p), _ptr(dp), rows, cols, ld, float(scale),
                                            1 if self._clean else 0, self._stream()))
        return dp

    def colsum(self, x, out=None):
        """x [G, rows, cols] (row stride arbitrary, inner contiguous) -> out [G, cols] (strided ok)."""
        if x.dim() == 2:
            x = x.unsqueeze(0)
        G, rows, cols = x.shape
        assert x.stride(2) == 1 or cols == 1
        assert G == 1 or x.stride(0) == rows * x.stride(1)
        if out is None:
            out = self.empty(G, cols)
        assert out.shape == (G, cols) and (cols == 1 or out.stride(1) == 1)
        # few groups x many rows (the shared sinks of the meta-training step sum over all episodes):
        # one block per 32 columns would walk all rows serially.  Split the rows over `d` blocks
        # (d | rows), reduce the d partial sums in a second pass - same fixed order every time.
        if rows >= 2048 and G * ((cols + 31) // 32) < 296 and x.stride(1) == cols:
            d = next((k for k in range(min(256, rows // 64), 1, -1) if rows % k == 0), 1)
            if d > 1:
                part = self.colsum(x.reshape(G * d, rows // d, cols))
                return self.colsum(part.view(G, d, cols), out=out)
        _lib.check(self.lib.itn_colsum(_ptr(x), _ptr(out), G, rows, cols, x.stride(1),
                                       out.stride(0) if G > 1 else cols, self._stream()))
        return out

    # -------------------------------------------------------- element-wise
    def add(self, a, b, rnd=False):
        """a [G, n, ...] + b, where b is either one block broadcast over everything
        (a.numel() % b.numel() == 0) or [G, block] with one block per leading group of a."""
        assert a.is_contiguous()
        out = self.empty(a.shape)
        n = a.numel()
        grouped = b.dim() >= 2 and b.shape[0] == a.shape[0] and b.shape[0] > 1
        if grouped and (n != b.numel() or not b.is_contiguous()):
            G = a.shape[0]
            assert b[0].is_contiguous()          # blocks may sit G rows apart in a flat buffer
            b_elems, a_group, b_gs = b.numel() // G, n // G, b.stride(0)
        else:
            assert b.is_contiguous()
            b_elems, a_group, b_gs = b.numel(), n, 0
        assert n % b_elems == 0
        _lib.check(self.lib.itn_add(_ptr(a), _ptr(b), _ptr(out), n, b_elems, a_group, b_gs,
                                    1 if (rnd and self._clean) else 0, self._stream()))
        return out

    def dropout(self, x, key, residual=None, out=None):
        """out = (residual +) x * keep / (1 - p) with the counter-based mask of itn_dropout; key = (p, device seed
        tensor int64[1], site).  x: contiguous [..., cols] or a 2-D view with unit inner stride (rows = mask rows).
        out may be x (in place); default: a new tensor."""
        p_, seed, site = key
        assert not self._clean, "train()-mode dropout runs in the tf32x3 (fp32-accurate) mode only"
        assert seed.is_cuda and seed.dtype == torch.int64 and seed.numel() == 1
        if x.dim() == 2 and not x.is_contiguous():
            assert x.stride(1) == 1
            x2 = x
        else:
            assert x.is_contiguous()
            x2 = x.view(-1, x.shape[-1])
        rows, cols = x2.shape
        if out is None:
            out = self.empty(*x.shape)
        o2 = out if (out.dim() == 2 and not out.is_contiguous()) else out.view(-1, out.shape[-1])
        assert tuple(o2.shape) == (rows, cols) and o2.stride(1) == 1
        r2 = None
        if residual is not None:
            assert residual.is_contiguous() and residual.numel() == rows * cols
            r2 = residual.view(rows, cols)
        _lib.check(self.lib.itn_dropout(_ptr(x2), x2.stride(0), _ptr(r2), cols if r2 is not None else 0, _ptr(o2),
                                        o2.stride(0), rows, cols, 0, float(p_), _ptr(seed), int(site), self._stream()))
        return out

    def copy2d_(self, dst, src, rnd=False):
        """dst[r, c] = src[r, c] for 2-D views with unit inner stride."""
        assert dst.shape == src.shape and dst.dim() == 2
        assert (dst.stride(1) == 1 and src.stride(1) == 1) or dst.shape[1] == 1
        _lib.check(self.lib.itn_copy2d(_ptr(src), src.stride(0), _ptr(dst), dst.stride(0), dst.shape[0],
                                       dst.shape[1], 1 if (rnd and self._clean) else 0, self._stream()))
        return dst

    def transpose_(self, dst, src):
        """dst [G, C, R] = src [G, R, C] transposed per group (inner dims contiguous)."""
        G, R, Cc = src.shape
        assert dst.shape == (G, Cc, R) and src[0].is_contiguous() and dst[0].is_contiguous()
        _lib.check(self.lib.itn_transpose(_ptr(src), _ptr(dst), G, R, Cc, src.stride(0) if G > 1 else 0,
                                          dst.stride(0) if G > 1 else 0, self._stream()))
        return dst

    def round_tf32(self, x, out=None):
        """TF32-rounded copy of contiguous x (out may alias x).  In tf32x3 mode operands are used
        at full precision, so this is the identity (no launch)."""
        if not self._clean:
            return x
        assert x.is_contiguous()
        out = self.empty(x.shape) if out is None else out
        _lib.check(self.lib.itn_round_tf32(_ptr(x), _ptr(out), x.numel(), self._stream()))
        return out

    def sigmoid(self, x):
        assert x.is_contiguous()
        y = self.empty(x.shape)
        _lib.check(self.lib.itn_sigmoid_fwd(_ptr(x), _ptr(y), x.numel(), self._stream()))
        return y

    def sigmoid_bwd(self, dy, y):
        assert dy.is_contiguous() and y.is_contiguous()
        dx = self.empty(y.shape)
        _lib.check(self.lib.itn_sigmoid_bwd(_ptr(dy), _ptr(y), _ptr(dx), y.numel(), self._stream()))
        return dx

    def l2norm_fwd_bwd(self, x):
        """x [G, n] -> (||x_g|| [G], x_g/||x_g|| [G, n])."""
        assert x.is_contiguous() and x.dim() == 2
        G, n = x.shape
        loss, dx = self.empty(G), self.empty(G, n)
        _lib.check(self.lib.itn_l2norm_fwd_bwd(_ptr(x), _ptr(loss), _ptr(dx), G, n, self._stream()))
        return loss, dx

    def sgd_clip_update(self, theta, g, lr, clip=0.01, want_mask=False):
        """theta [n] (shared) or [G,n]; g [G,n] -> theta' [G,n], TF32-rounded theta' (, uint8 mask)."""
        assert g.is_contiguous() and g.dim() == 2 and theta.is_contiguous()
        G, n = g.shape
        stride = 0 if theta.dim() == 1 or theta.shape[0] == 1 else n
        out = self.empty(G, n)
        out_r = self.empty(G, n) if self._clean else None
        mask = torch.empty(G, n, dtype=torch.uint8, device=self.device) if want_mask else None
        _lib.check(self.lib.itn_sgd_clip_update(_ptr(theta), stride, _ptr(g), _ptr(out), _ptr(out_r), _ptr(mask),
                                                G, n, float(lr), float(clip), self._stream()))
        out_r = out if out_r is None else out_r
        return (out, out_r, mask) if want_mask else (out, out_r)

    # ---------------------------------------- tangent rules (meta-training step, dual.py)
    @staticmethod
    def _gb(t, cols):
        """[G, cols] view of an affine parameter (row stride arbitrary) -> (tensor, groups, stride)."""
        if t is None:
            return None, 1, 0
        t2 = t.reshape(-1, cols)
        assert t2.stride(1) == 1
        return t2, t2.shape[0], (t2.stride(0) if t2.shape[0] > 1 else 0)

    def layernorm_fwd_jvp(self, x, x_dot, mean, rstd, gamma, gamma_dot, beta_dot):
        rows, cols = x.shape
        assert x.is_contiguous() and (x_dot is None or x_dot.is_contiguous())
        g2, G, gs = self._gb(gamma, cols)
        gd, Gd, ds = self._gb(gamma_dot, cols)
        bd, Gb, bs = self._gb(beta_dot, cols)
        if gd is not None and bd is not None:
            assert Gd == Gb and ds == bs
        elif bd is not None:
            Gd, ds = Gb, bs
        y_dot = self.empty(rows, cols)
        _lib.check(self.lib.itn_layernorm_fwd_jvp(_ptr(x), _ptr(x_dot), _ptr(mean), _ptr(rstd), _ptr(g2), _ptr(gd),
                                                  _ptr(bd), _ptr(y_dot), rows, cols, G, gs, Gd, ds, self._stream()))
        return y_dot

    def layernorm_bwd_jvp(self, dy, dy_dot, x, x_dot, mean, rstd, gamma, gamma_dot, dgamma_dot=None, dbeta_dot=None):
        rows, cols = x.shape
        assert dy.is_contiguous() and x.is_contiguous()
        assert (dy_dot is None or dy_dot.is_contiguous()) and (x_dot is None or x_dot.is_contiguous())
        g2, G, gs = self._gb(gamma, cols)
        gd, Gd, ds = self._gb(gamma_dot, cols)
        dx_dot = self.empty(rows, cols)
        gterm = self.empty(rows, cols) if dgamma_dot is not None else None
        _lib.check(self.lib.itn_layernorm_bwd_jvp(_ptr(dy), _ptr(dy_dot), _ptr(x), _ptr(x_dot), _ptr(mean), _ptr(rstd),
                                                  _ptr(g2), _ptr(gd), _ptr(dx_dot), _ptr(gterm), rows, cols, G, gs,
                                                  Gd, ds, self._stream()))
        if dgamma_dot is not None:
            Gn = dgamma_dot.shape[0]
            self.colsum(gterm.view(Gn, rows // Gn, cols), out=dgamma_dot)
            if dy_dot is not None:
                self.colsum(dy_dot.view(Gn, rows // Gn, cols), out=dbeta_dot)
        return dx_dot

    def softmax_bwd_jvp_(self, p, p_dot, dp, dp_dot, cols, scale):
        assert p.is_contiguous() and dp.is_contiguous() and dp_dot.is_contiguous() and p.shape == dp.shape
        assert p_dot is None or (p_dot.is_contiguous() and p_dot.shape == p.shape)
        ld = p.shape[-1]
        rows = p.numel() // ld
        _lib.check(self.lib.itn_softmax_bwd_jvp(_ptr(p), _ptr(p_dot), _ptr(dp), _ptr(dp_dot), rows, cols, ld,
                                                float(scale), self._stream()))
        return dp

    def mask_mul_(self, y, ref):
        assert y.is_contiguous() and ref.is_contiguous() and y.numel() == ref.numel()
        _lib.check(self.lib.itn_mask_mul(_ptr(y), _ptr(ref), y.numel(), self._stream()))
        return y

    def mul_mask_u8(self, x, mask, scale=1.0):
        assert x.is_contiguous() and mask.is_contiguous() and mask.dtype == torch.uint8 and mask.numel() == x.numel()
        out = self.empty(x.shape)
        _lib.check(self.lib.itn_mul_mask_u8(_ptr(x), _ptr(mask), float(scale), _ptr(out), x.numel(), self._stream()))
        return out

    def gelu_grad_dual(self, raw, raw_dot, aux, aux_dot, out=None):
        assert raw.is_contiguous() and aux.is_contiguous() and raw.numel() == aux.numel()
        for t in (raw_dot, aux_dot, out):
            assert t is None or (t.is_contiguous() and t.numel() == raw.numel())
        y = self.empty(raw.shape) if out is None else out
        y_dot = self.empty(raw.shape) if (raw_dot is not None or aux_dot is not None) else None
        _lib.check(self.lib.itn_gelu_grad_dual(_ptr(raw), _ptr(raw_dot), _ptr(aux), _ptr(aux_dot), _ptr(y), _ptr(y_dot),
                                               raw.numel(), self._stream()))
        return y, y_dot

    def sigmoid_bwd_jvp(self, dy, dy_dot, y, y_dot):
        for t in (dy, dy_dot, y, y_dot):
            assert t is None or t.is_contiguous()
        out = self.empty(y.shape)
        _lib.check(self.lib.itn_sigmoid_bwd_jvp(_ptr(dy), _ptr(dy_dot), _ptr(y), _ptr(y_dot), _ptr(out), y.numel(),
                                                self._stream()))
        return out

    def l2norm_jvp(self, x_dot, nrm, d):
        assert x_dot.is_contiguous() and d.is_contiguous() and x_dot.dim() == 2
        G, n = x_dot.shape
        n_dot, d_dot = self.empty(G), self.empty(G, n)
        _lib.check(self.lib.itn_l2norm_jvp(_ptr(x_dot), _ptr(nrm), _ptr(d), _ptr(n_dot), _ptr(d_dot), G, n,
                                           self._stream()))
        return n_dot, d_dot

    def im2col_nhwc(self, x, kh, kw, stride, pad, dil):
        """x [N,H,W,C] channels-last -> ([N*Ho*Wo, ld] patch matrix, Ho, Wo); ld = kh*kw*C rounded up to 4."""
        assert x.is_contiguous() and x.dim() == 4
        N, H, W, Cc = x.shape
        Ho = (H + 2 * pad - dil * (kh - 1) - 1) // stride + 1
        Wo = (W + 2 * pad - dil * (kw - 1) - 1) // stride + 1
        ld = (kh * kw * Cc + 3) // 4 * 4
        out = self.empty(N * Ho * Wo, ld)
        _lib.check(self.lib.itn_im2col_nhwc(_ptr(x), _ptr(out), N, H, W, Cc, kh, kw, stride, pad, dil, Ho, Wo, ld,
                                            self._stream()))
        return out, Ho, Wo

    # ------------------------------------------------------------ trainer step
    def sumsq_partials(self, g):
        """Fixed-order partial sums of squares of the flat gradient buffer -> float32 [n_partials]."""
        assert g.is_contiguous()
        part = self.empty(1184)
        n = C.c_int(0)
        _lib.check(self.lib.itn_sumsq_partials(_ptr(g), g.numel(), _ptr(part), part.numel(), C.byref(n), self._stream()))
        return part[:n.value]

    def clip_adam_step_(self, w, g, m, v, partials, max_norm, lr, betas, eps, step, zero_grad=False, norm_out=None):
        """In place on one flat segment: g *= clip coefficient of the global norm sqrt(sum(partials)),
        Adam moments and weights updated as torch.optim.Adam does at `step` (1-based)."""
        for t in (w, g, m, v):
            assert t.is_contiguous() and t.numel() == w.numel() and t.dtype == torch.float32
        _lib.check(self.lib.itn_clip_adam_step(
            _ptr(w), _ptr(g), _ptr(m), _ptr(v), w.numel(), _ptr(partials), 0 if partials is None else partials.numel(),
            float(max_norm), float(lr), float(betas[0]), float(betas[1]), float(eps), int(step), 1 if zero_grad else 0,
            _ptr(norm_out), self._stream()))

    def ckpt_accumulate_(self, acc, x, w, first):
        """acc = (first ? 0 : acc) + w * x on flat buffers, in torch's operation order (trainer.CheckpointAverager)."""
        assert acc.is_contiguous() and x.is_contiguous() and acc.numel() == x.numel() and acc.dtype == torch.float32
        _lib.check(self.lib.itn_ckpt_accumulate(_ptr(x), _ptr(acc), x.numel(), float(w), 1 if first else 0,
                                                self._stream()))
        return acc

    # ------------------------------------------------------------ evaluator post-processing
    def detect_postprocess(self, logits, boxes, background, iou_threshold=0.5):
        """logits [I,Q,C], boxes [I,Q,4] cxcywh -> (count [I] int32, keep_idx [I,Q] int32, score [I,Q],
        cat [I,Q] int32, xyxy [I,Q,4]): softmax-max, background filter, class-agnostic NMS per image."""
        assert logits.is_contiguous() and boxes.is_contiguous() and logits.dim() == 3
        I, Q, Cc = logits.shape
        i32 = lambda *s: torch.empty(*s, dtype=torch.int32, device=self.device)
        count, keep, cat = i32(I), i32(I, Q), i32(I, Q)
        score, xyxy = self.empty(I, Q), self.empty(I, Q, 4)
        _lib.check(self.lib.itn_detect_postprocess(_ptr(logits), _ptr(boxes), I, Q, Cc, int(background),
                                                   float(iou_threshold), _ptr(count), _ptr(keep), _ptr(score),
                                                   _ptr(cat), _ptr(xyxy), self._stream()))
        return count, keep, score, cat, xyxy

    def maxpool3x3s2_nhwc(self, x):
        assert x.is_contiguous() and x.dim() == 4
        N, H, W, Cc = x.shape
        Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        out = self.empty(N, Ho, Wo, Cc)
        _lib.check(self.lib.itn_maxpool3x3s2_nhwc(_ptr(x), _ptr(out), N, H, W, Cc, Ho, Wo, self._stream()))
        return out

    def pos_embed_sine(self, mask, feats=128):
        """mask [F,h,w] bool/uint8 (1 = padded) -> [F, h*w, 2*feats]."""
        m = mask.to(torch.uint8).contiguous()
        F_, h, w = m.shape
        pos = self.empty(F_, h * w, 2 * feats)
        _lib.check(self.lib.itn_pos_embed_sine(_ptr(m), _ptr(pos), F_, h, w, feats, self._stream()))
        return pos

    def matcher_cost(self, logits, boxes, tgt_boxes, tgt_labels, tgt_off, w_class, w_bbox, w_giou):
        """Block-diagonal HungarianMatcher cost, packed per frame ([Q, T_f] row-major)."""
        F_, Q, Cn = logits.shape
        assert logits.is_contiguous() and boxes.is_contiguous() and tgt_boxes.is_contiguous()
        total = int(tgt_boxes.shape[0])
        cost = self.empty(max(Q * total, 1))
        _lib.check(self.lib.itn_matcher_cost(_ptr(logits), _ptr(boxes), _ptr(tgt_boxes), _ptr(tgt_labels),
                                             _ptr(tgt_off), _ptr(cost), F_, Q, Cn, float(w_class),
                                             float(w_bbox), float(w_giou), self._stream()))
        return cost

    def criterion(self, logits, boxes, tgt_boxes, tgt_labels, tgt_off, match_row, match_tgt, match_off,
                  groups, background_c, weights=(1.0, 1.0, 1.0), want_grad=False):
        """SetCriterion on the device (see itn_criterion): logits [groups*F,Q,C], boxes [groups*F,Q,4].
        -> losses [groups,5] (loss_ce, class_error, cardinality_error, loss_bbox, loss_giou)
        and, with want_grad, d(w_ce*ce + w_bbox*bbox + w_giou*giou)/d(logits, boxes)."""
        Fn, Q, Cn = logits.shape
        assert Fn % groups == 0 and logits.is_contiguous() and boxes.is_contiguous()
        n_match = int(match_row.numel())
        rows = Fn * Q
        nbytes = self.lib.itn_criterion_scratch_bytes(rows, n_match, groups)
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        losses = self.empty(groups, 5)
        dl = self.empty(Fn, Q, Cn) if want_grad else None
        db = self.empty(Fn, Q, 4) if want_grad else None
        _lib.check(self.lib.itn_criterion(
            _ptr(logits), _ptr(boxes), _ptr(tgt_boxes) if n_match else None, _ptr(tgt_labels) if n_match else None,
            _ptr(tgt_off), _ptr(match_row) if n_match else None, _ptr(match_tgt) if n_match else None,
            _ptr(match_off), n_match, groups, Fn // groups, Q, Cn, float(background_c), float(weights[0]),
            float(weights[1]), float(weights[2]), _ptr(losses), _ptr(dl) if want_grad else None,
            _ptr(db) if want_grad else None, _ptr(scratch), self._stream()))
        return (losses, dl, db) if want_grad else losses
