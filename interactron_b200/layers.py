"""Building blocks shared by the DETR transformer and the fusion networks: attention core,
post-norm decoder layer (forward + hand-derived backward), MLP heads, gradient sink.

Everything is written against the `ops` kernel interface (ops.CudaOps on the B200).  Activations
are token-major 2-D/3-D tensors; attention heads are strided views of [batch, tokens, heads*hd].
Tensors that feed a GEMM are stored TF32-rounded by their producer (rnd=True / `_r` twins);
residual streams and gradient accumulators stay full fp32.
"""
import torch


def pad4(n):
    return (n + 3) // 4 * 4


def T(w):
    return w.transpose(-1, -2)


def lin(ops, x, W, b=None, **kw):
    """x [E,R,K] @ W[Gw,N,K]^T (+ b[Gw,N]) -> [E,R,N]  (nn.Linear with grouped weights)."""
    return ops.matmul(x, T(W), bias=b, **kw)


# --------------------------------------------------------------------------- train()-mode dropout
class DropCtx:
    """Dropout state of ONE forward pass in train() mode (the reference trainers call model.train():
    engine/interactron_trainer.py:73): probability, the device seed tensor (int64[1], rewritten by the host before
    every step so that replayed CUDA graphs draw fresh masks) and the running site counter.  `next()` is called
    once per dropout module in the reference's forward order (models/detr_models/transformer.py:154-159,219-230;
    models/gpt.py:51,56,72,195); the key it returns is kept in the layer cache so that the backward pass and the
    dual-number pass regenerate the same mask (csrc/itn_philox.cuh)."""

    def __init__(self, p, seed, first_site=0, p_attn=None, p_embd=None):
        self.p, self.seed, self.site = float(p), seed, int(first_site)
        self.p_kind = {"attn": float(p if p_attn is None else p_attn), "embd": float(p if p_embd is None else p_embd)}

    def next(self, kind=None):
        """kind: None (sub-layer / FFN dropout), "attn" (attention probabilities), "embd" (GPT embedding dropout):
        fusion A has one probability per kind (ATTENTION_PDROP / RESIDUAL_PDROP / EMBEDDING_PDROP)."""
        key = (self.p_kind.get(kind, self.p), self.seed, self.site)
        self.site += 1
        return key


def _next(drop, kind=None):
    return None if drop is None else drop.next(kind)


def _drop_res(ops, y, key, residual):
    """residual + dropout(y): what `x + self.dropoutN(y)` computes in train() mode."""
    return ops.dropout(y, key, residual=residual)


# --------------------------------------------------------------------------- attention core
def attention_fwd(ops, q, k, v, B, Lq, Lk, nh, hd, scale, kmask, drop=None):
    """softmax(scale * q k^T + mask) v per (batch, head).
    q [B,Lq,nh*hd], k/v [B,Lk,nh*hd] (strided views ok) -> o [B,Lq,nh*hd], ctx (what attention_bwd needs).
    On the B200 backend this is ONE fused tcgen05 kernel (itn_attention_fwd: the score matrix stays in
    tensor memory, ctx = (o, log-sum-exp)); backends without it (the float64 CPU simulation, the
    dual-number pass of the meta-training step) run the unfused chain and keep the probabilities."""
    fused = getattr(ops, "attention_supported", None)
    if fused is not None and fused(q, k, v, nh):
        o, lse = ops.attention_fwd(q, k, v, nh, scale, kmask, drop=drop)
        return o, FusedCtx(o, lse, kmask, drop)
    return _attention_fwd_unfused(ops, q, k, v, B, Lq, Lk, nh, hd, scale, kmask, drop)


class FusedCtx:
    """Saved state of a fused attention forward: output and base-2 log-sum-exp per row."""

    __slots__ = ("o", "lse", "kmask", "drop")

    def __init__(self, o, lse, kmask, drop=None):
        self.o, self.lse, self.kmask, self.drop = o, lse, kmask, drop


class UnfusedCtx:
    """Saved state of an unfused attention forward in train() mode: probabilities before and after dropout."""

    __slots__ = ("P", "Pd", "drop")

    def __init__(self, P, Pd, drop):
        self.P, self.Pd, self.drop = P, Pd, drop


def _attention_fwd_unfused(ops, q, k, v, B, Lq, Lk, nh, hd, scale, kmask, drop=None):
    """QK^T GEMM -> softmax -> PV GEMM with the probabilities P [B,nh,Lq,pad4(Lk)] materialised."""
    P = ops.empty(B, nh, Lq, pad4(Lk))                           # row stride padded to 16 bytes for TMA
    o = ops.empty(B, Lq, nh * hd)
    if drop is not None:
        # attention-probability dropout (nn.MultiheadAttention(dropout=0.1), gpt.py:51): mask row = (b*nh+h)*Lq+i
        qh = q.reshape(B, Lq, nh, hd).permute(0, 2, 1, 3)
        khT = k.reshape(B, Lk, nh, hd).permute(0, 2, 3, 1)
        vh = v.reshape(B, Lk, nh, hd).permute(0, 2, 1, 3)
        ops.matmul(qh, khT, out=P[..., :Lk], out_pad=True)
        ops.softmax_(P, Lk, scale, kmask, rows_per_mask=nh * Lq)
        Pd = ops.zeros(B, nh, Lq, pad4(Lk))
        ops.dropout(P.view(B * nh * Lq, pad4(Lk))[:, :Lk], drop, out=Pd.view(B * nh * Lq, pad4(Lk))[:, :Lk])
        ops.matmul(Pd[..., :Lk], vh, out=o.view(B, Lq, nh, hd).permute(0, 2, 1, 3), rnd=True)
        return o, UnfusedCtx(P, Pd, drop)
    for b0, b1 in _l2_chunks(ops, B, nh * Lq * pad4(Lk) * 4, 1):
        qh = q[b0:b1].reshape(b1 - b0, Lq, nh, hd).permute(0, 2, 1, 3)            # [b,nh,Lq,hd]
        khT = k[b0:b1].reshape(b1 - b0, Lk, nh, hd).permute(0, 2, 3, 1)           # [b,nh,hd,Lk]
        vh = v[b0:b1].reshape(b1 - b0, Lk, nh, hd).permute(0, 2, 1, 3)            # [b,nh,Lk,hd]
        Pc = P[b0:b1]
        p = Pc[..., :Lk]
        ops.matmul(qh, khT, out=p, out_pad=True)
        ops.softmax_(Pc, Lk, scale, None if kmask is None else kmask[b0:b1], rows_per_mask=nh * Lq)
        ops.matmul(p, vh, out=o[b0:b1].view(b1 - b0, Lq, nh, hd).permute(0, 2, 1, 3), rnd=True)
    return o, P


def _l2_chunks(ops, B, bytes_per_batch, live):
    """Batch ranges whose score tensors (`live` of them, `bytes_per_batch` each per batch entry) fit the
    L2 budget `ops.attn_l2_mb`: the unfused QK^T -> softmax -> PV chain then re-reads its score tile
    from L2 instead of HBM (same kernels on the same rows: results are bit-identical to one pass).
    Dual-number passes (Dual tensors do not slice) and small problems take one chunk."""
    budget = getattr(ops, "attn_l2_mb", 0) * (1 << 20)
    total = B * bytes_per_batch * live
    if budget <= 0 or total <= budget:
        return [(0, B)]
    step = max(1, int(budget // (bytes_per_batch * live)))
    n = -(-B // step)
    step = -(-B // n)                                            # equal-sized chunks
    return [(b, min(B, b + step)) for b in range(0, B, step)]


def attention_bwd(ops, dO, q, k, v, P, B, Lq, Lk, nh, hd, scale, dq, dk, dv):
    """dO [B,Lq,nh*hd] (TF32-clean).  Writes TF32-clean dq/dk/dv into the given [B,L,nh*hd] views.
    P is the ctx attention_fwd returned: FusedCtx (scores recomputed on chip) or the probabilities."""
    if isinstance(P, FusedCtx):
        ops.attention_bwd(dO, q, k, v, P.o, P.lse, nh, scale, P.kmask, dq, dk, dv, drop=P.drop)
        return
    if isinstance(P, UnfusedCtx):
        ctx, ld = P, pad4(Lk)
        qh = q.reshape(B, Lq, nh, hd).permute(0, 2, 1, 3)
        kh = k.reshape(B, Lk, nh, hd).permute(0, 2, 1, 3)
        vhT = v.reshape(B, Lk, nh, hd).permute(0, 2, 3, 1)
        dOh = dO.reshape(B, Lq, nh, hd).permute(0, 2, 1, 3)
        dP = ops.zeros(B, nh, Lq, ld)
        dp = dP[..., :Lk]
        ops.matmul(dOh, vhT, out=dp, out_pad=True)                                              # d(dropped P)
        ops.matmul(T(ctx.Pd[..., :Lk]), dOh, out=dv.view(B, Lk, nh, hd).permute(0, 2, 1, 3), rnd=True)
        ops.dropout(dP.view(B * nh * Lq, ld)[:, :Lk], ctx.drop, out=dP.view(B * nh * Lq, ld)[:, :Lk])    # -> dP
        ops.softmax_bwd_(ctx.P, dP, Lk, scale)
        ops.matmul(dp, kh, out=dq.view(B, Lq, nh, hd).permute(0, 2, 1, 3), rnd=True)
        ops.matmul(T(dp), qh, out=dk.view(B, Lk, nh, hd).permute(0, 2, 1, 3), rnd=True)
        return
    chunks = _l2_chunks(ops, B, nh * Lq * pad4(Lk) * 4, 2)
    # one chunk: dP for the whole batch; several: ONE chunk-sized dP buffer reused (it stays in L2)
    dP_all = ops.empty(chunks[0][1] - chunks[0][0], nh, Lq, pad4(Lk))
    for b0, b1 in chunks:
        n = b1 - b0
        qh = q[b0:b1].reshape(n, Lq, nh, hd).permute(0, 2, 1, 3)
        kh = k[b0:b1].reshape(n, Lk, nh, hd).permute(0, 2, 1, 3)
        vhT = v[b0:b1].reshape(n, Lk, nh, hd).permute(0, 2, 3, 1)
        dOh = dO[b0:b1].reshape(n, Lq, nh, hd).permute(0, 2, 1, 3)
        Pc = P[b0:b1]
        p = Pc[..., :Lk]
        dP = dP_all[:n]
        dp = dP[..., :Lk]
        ops.matmul(dOh, vhT, out=dp, out_pad=True)                                                    # dP = dO V^T
        ops.matmul(T(p), dOh, out=dv[b0:b1].view(n, Lk, nh, hd).permute(0, 2, 1, 3), rnd=True)  # dV = P^T dO
        ops.softmax_bwd_(Pc, dP, Lk, scale)                                              # dS (in dP)
        ops.matmul(dp, kh, out=dq[b0:b1].view(n, Lq, nh, hd).permute(0, 2, 1, 3), rnd=True)     # dQ = dS K
        ops.matmul(T(dp), qh, out=dk[b0:b1].view(n, Lk, nh, hd).permute(0, 2, 1, 3), rnd=True)  # dK = dS^T Q


# --------------------------------------------------------------------------- gradient sinks
class NullSink:
    """Data gradients only (the inner loop never needs d/d phi of the fusion network)."""

    def wants(self, name):
        return False

    def linear(self, name, dy_r, x_r, dy_full=None, accumulate=False):
        pass

    def rows(self, name_w, name_b, lo, hi, dy, x):
        pass

    def norm(self, name, bias_of=None):
        return {}

    def colsum(self, name, x):
        pass

    def copy(self, name, x):
        pass


class GradSink(NullSink):
    """Writes weight gradients straight into a flat gradient buffer laid out by `pack`.

    per-episode (default): g [E, n], one gradient per episode - the inner gradient of the fast
    weights theta (reference models/interactron.py:51-52).
    shared=True: g [1, n], the gradient summed over the episodes of the step - what `.backward()`
    accumulates on parameters every episode shares (in_proj_* and the fusion network in the
    meta-training step, reference models/interactron.py:123,134).  Activations [E, rows, .] are
    then read as one [1, E*rows, .] matrix.
    Names outside `pack` are ignored, so sinks over disjoint packs compose with MultiSink."""

    def __init__(self, ops, pack, g, shared=False):
        self.ops, self.pack, self.g, self.shared = ops, pack, g, shared
        self._bias_done = set()          # biases whose gradient came out of a fused LayerNorm backward

    def wants(self, name):
        return name in self.pack

    def view(self, name):
        return self.pack.view(self.g, name)

    def _flat(self, t):
        return t.reshape(1, -1, t.shape[-1]) if self.shared else t

    def linear(self, name, dy_r, x_r, dy_full=None, accumulate=False):
        """dW = dy^T x into `name.weight`, db = colsum(dy) into `name.bias` (if there is one).
        accumulate: second use of a bias-free weight in the same pass (fusion A's `model.head`)."""
        if not self.wants(name + ".weight"):
            return
        w = self.view(name + ".weight")
        self.ops.matmul(T(self._flat(dy_r)), self._flat(x_r), out=w.reshape(w.shape[0], w.shape[1], -1),
                        accumulate=accumulate)
        if self.wants(name + ".bias"):
            assert not accumulate
            if name in self._bias_done:
                self._bias_done.discard(name)
                return
            self.ops.colsum(self._flat(dy_full if dy_full is not None else dy_r), out=self.view(name + ".bias"))

    def rows(self, name_w, name_b, lo, hi, dy, x):
        """Rows [lo, hi) of a packed projection (nn.MultiheadAttention.in_proj_*)."""
        if not self.wants(name_w):
            return
        self.ops.matmul(T(self._flat(dy)), self._flat(x), out=self.view(name_w)[:, lo:hi])
        self.ops.colsum(self._flat(dy), out=self.view(name_b)[:, lo:hi])

    def norm(self, name, bias_of=None):
        """Keyword arguments for ops.layernorm_bwd.  bias_of: the linear layer whose output is the residual
        branch this LayerNorm normalises (post-norm block, no dropout in between): its bias gradient is
        colsum(dx), which the fused LayerNorm backward of the B200 backend produces in the same launch
        (`dxsum`); `linear(bias_of, ...)` then skips its column sum."""
        if not self.wants(name + ".weight"):
            return {}
        kw = dict(dgamma=self.view(name + ".weight"), dbeta=self.view(name + ".bias"))
        if bias_of is not None and getattr(self.ops, "fused_ln_bwd", False) and self.wants(bias_of + ".bias"):
            b = self.view(bias_of + ".bias")
            kw["dxsum"] = b.reshape(b.shape[0], -1)
            self._bias_done.add(bias_of)
        return kw

    def colsum(self, name, x):
        """x [G, rows, cols] summed over rows into parameter `name` (numel = cols per group)."""
        if not self.wants(name):
            return
        if self.shared:
            x = x.reshape(1, -1, x.shape[-1])
        v = self.view(name)
        self.ops.colsum(x, out=v.reshape(v.shape[0], -1))

    def copy(self, name, x):
        """x [G, numel] is the gradient of parameter `name` itself."""
        if not self.wants(name):
            return
        v = self.view(name)
        self.ops.copy2d_(v.reshape(v.shape[0], -1), x)


class MultiSink(NullSink):
    def __init__(self, *sinks):
        self.sinks = [s for s in sinks if s is not None]

    def wants(self, name):
        return any(s.wants(name) for s in self.sinks)

    def linear(self, name, dy_r, x_r, dy_full=None, accumulate=False):
        for s in self.sinks:
            s.linear(name, dy_r, x_r, dy_full, accumulate)

    def rows(self, name_w, name_b, lo, hi, dy, x):
        for s in self.sinks:
            s.rows(name_w, name_b, lo, hi, dy, x)

    def norm(self, name, bias_of=None):
        for s in self.sinks:
            if s.wants(name + ".weight"):
                return s.norm(name, bias_of)
        return {}

    def colsum(self, name, x):
        for s in self.sinks:
            s.colsum(name, x)

    def copy(self, name, x):
        for s in self.sinks:
            s.copy(name, x)


# --------------------------------------------------------------------------- decoder layer
class DecDims:
    """E: episodes (weight groups of the activations), B: attention batches, Lq/Lk tokens per batch."""

    def __init__(self, E, B, Lq, Lk, D, nh):
        self.E, self.B, self.Lq, self.Lk, self.D, self.nh = E, B, Lq, Lk, D, nh
        self.hd = D // nh
        self.scale = 1.0 / (self.hd ** 0.5)
        self.Q = B * Lq // E           # query rows per episode
        self.R = B * Lk // E           # memory rows per episode


def ln_plus(ops, x, gamma, beta, plus, shape):
    """LayerNorm forward + `y + plus` (the q/k input of the next attention): one launch where the backend has
    layernorm_fwd_plus (B200), LayerNorm then add elsewhere.  -> y, y_r, mean, rstd, y_plus [rows, cols]."""
    f = getattr(ops, "layernorm_fwd_plus", None)
    if f is not None:
        return f(x, gamma, beta, plus, shape)
    y, y_r, mean, rstd = ops.layernorm_fwd(x, gamma, beta)
    return y, y_r, mean, rstd, ops.add(y.view(shape), plus, rnd=True).view(x.shape)


def _stacked(views):
    """Equally spaced, identically shaped views of one buffer -> one view with a new dim 1 (layer) inserted
    after the group dim; None if the views do not line up (or are not plain tensors, e.g. dual numbers)."""
    v0 = views[0]
    if not all(isinstance(v, torch.Tensor) for v in views) or len(views) < 2:
        return None
    base = v0.untyped_storage().data_ptr()
    step = views[1].storage_offset() - v0.storage_offset()
    for j, v in enumerate(views):
        if (v.untyped_storage().data_ptr() != base or v.shape != v0.shape or v.stride() != v0.stride()
                or v.dtype != v0.dtype or v.storage_offset() != v0.storage_offset() + j * step):
            return None
    if step <= 0:
        return None
    return torch.as_strided(v0, (v0.shape[0], len(views)) + tuple(v0.shape[1:]),
                            (v0.stride(0), step) + tuple(v0.stride()[1:]), v0.storage_offset())


def cross_kv_all(ops, W, pres, dm, mem_pos_r, memory_r):
    """Cross-attention key / value projections of ALL decoder layers `pres` at once: they depend on the encoder
    memory only (reference detr_models/transformer.py:222-226 recomputes them inside every layer), so the n
    layers' K projections are ONE GEMM batched over the layer dim (and one for V) when the layers'
    in_proj tensors are equally spaced views of one flat parameter buffer (they are: identical layers packed
    back to back).  -> [(kc_j, vc_j)] per layer ([B, Lk, D] each), or None (caller projects per layer)."""
    D, B, Lk = dm.D, dm.B, dm.Lk
    ws = [W.w(p + "multihead_attn.in_proj_weight") for p in pres]
    bs = [W.p(p + "multihead_attn.in_proj_bias") for p in pres]
    if not isinstance(mem_pos_r, torch.Tensor) or ws[0].shape[0] != 1:
        return None
    wk, wv = _stacked([w[:, D:2 * D] for w in ws]), _stacked([w[:, 2 * D:] for w in ws])
    bk, bv = _stacked([b[:, D:2 * D] for b in bs]), _stacked([b[:, 2 * D:] for b in bs])
    if wk is None or wv is None or bk is None or bv is None:
        return None
    kc = ops.matmul(mem_pos_r.unsqueeze(1), wk.transpose(-1, -2), bias=bk, rnd=True)       # [1, n, E*R, D]
    vc = ops.matmul(memory_r.unsqueeze(1), wv.transpose(-1, -2), bias=bv, rnd=True)
    return [(kc[0, j].view(B, Lk, D), vc[0, j].view(B, Lk, D)) for j in range(len(pres))]


def decoder_layer_fwd(ops, W, pre, dm, tgt, tgt_r, qpos, mem_pos_r, memory_r, kmask, need_cache=True, drop=None,
                      tgt_plus=None, want_plus=False, kv=None):
    """DETR post-norm decoder layer (reference detr_models/transformer.py:211-232).
    tgt/tgt_r [E,Q,D]; qpos [Gw,Lq,D] added to the queries of every batch; mem_pos_r / memory_r
    [1,E*R,D] TF32-clean keys-input (memory+pos) and values-input (memory).  -> t3, t3_r, cache.
    tgt_plus: tgt + qpos if the caller already has it (the previous layer's want_plus output);
    want_plus: also return t3 + qpos (fused into the last LayerNorm) as a fourth value.
    kv: this layer's (kc, vc) from cross_kv_all (projected for all layers at once), else projected here."""
    E, B, Lq, Lk, D, nh, hd, Q, R = dm.E, dm.B, dm.Lq, dm.Lk, dm.D, dm.nh, dm.hd, dm.Q, dm.R
    sw, sb = W.w(pre + "self_attn.in_proj_weight"), W.p(pre + "self_attn.in_proj_bias")
    cw, cb = W.w(pre + "multihead_attn.in_proj_weight"), W.p(pre + "multihead_attn.in_proj_bias")
    # self attention among the Lq queries of each batch
    qk_in = (tgt_plus if tgt_plus is not None else ops.add(tgt, qpos, rnd=True)).view(1, E * Q, D)
    qk = lin(ops, qk_in, sw[:, :2 * D], sb[:, :2 * D], rnd=True)
    v = lin(ops, tgt_r.view(1, E * Q, D), sw[:, 2 * D:], sb[:, 2 * D:], rnd=True)
    qk3, v3 = qk.view(B, Lq, 2 * D), v.view(B, Lq, D)
    # train(): dropout keys in the reference's module order (transformer.py:219-230): self-attention
    # probabilities, dropout1, cross-attention probabilities, dropout2, FFN dropout, dropout3
    ka1 = _next(drop, "attn")
    o, P = attention_fwd(ops, qk3[..., :D], qk3[..., D:], v3, B, Lq, Lq, nh, hd, dm.scale, None, drop=ka1)
    kd1 = _next(drop)
    if drop is None:
        a1 = lin(ops, o.view(E, Q, D), W.w(pre + "self_attn.out_proj.weight"),
                 W.p(pre + "self_attn.out_proj.bias"), residual=tgt)
    else:
        a1 = _drop_res(ops, lin(ops, o.view(E, Q, D), W.w(pre + "self_attn.out_proj.weight"),
                                W.p(pre + "self_attn.out_proj.bias")), kd1, tgt)
    t1, t1_r, m1, r1, q_in = ln_plus(ops, a1.view(E * Q, D), W.p(pre + "norm1.weight"), W.p(pre + "norm1.bias"),
                                     qpos, (E, Q, D))
    # cross attention into the Lk memory tokens of the batch
    q_in = q_in.view(1, E * Q, D)
    qc = lin(ops, q_in, cw[:, :D], cb[:, :D], rnd=True).view(B, Lq, D)
    if kv is not None:
        kc, vc = kv
    else:
        kc = lin(ops, mem_pos_r, cw[:, D:2 * D], cb[:, D:2 * D], rnd=True).view(B, Lk, D)
        vc = lin(ops, memory_r, cw[:, 2 * D:], cb[:, 2 * D:], rnd=True).view(B, Lk, D)
    ka2 = _next(drop, "attn")
    o2, P2 = attention_fwd(ops, qc, kc, vc, B, Lq, Lk, nh, hd, dm.scale, kmask, drop=ka2)
    kd2 = _next(drop)
    if drop is None:
        a2 = lin(ops, o2.view(E, Q, D), W.w(pre + "multihead_attn.out_proj.weight"),
                 W.p(pre + "multihead_attn.out_proj.bias"), residual=t1.view(E, Q, D))
    else:
        a2 = _drop_res(ops, lin(ops, o2.view(E, Q, D), W.w(pre + "multihead_attn.out_proj.weight"),
                                W.p(pre + "multihead_attn.out_proj.bias")), kd2, t1.view(E, Q, D))
    t2, t2_r, m2, r2 = ops.layernorm_fwd(a2.view(E * Q, D), W.p(pre + "norm2.weight"), W.p(pre + "norm2.bias"))
    h = lin(ops, t2_r.view(E, Q, D), W.w(pre + "linear1.weight"), W.p(pre + "linear1.bias"), act="relu", rnd=True)
    kf, kd3 = _next(drop), _next(drop)
    if drop is None:
        f = lin(ops, h, W.w(pre + "linear2.weight"), W.p(pre + "linear2.bias"), residual=t2.view(E, Q, D))
    else:
        h = ops.dropout(h, kf, out=h)
        f = _drop_res(ops, lin(ops, h, W.w(pre + "linear2.weight"), W.p(pre + "linear2.bias")), kd3, t2.view(E, Q, D))
    t3_plus = None
    if want_plus:
        t3, t3_r, m3, r3, t3_plus = ln_plus(ops, f.view(E * Q, D), W.p(pre + "norm3.weight"), W.p(pre + "norm3.bias"),
                                            qpos, (E, Q, D))
    else:
        t3, t3_r, m3, r3 = ops.layernorm_fwd(f.view(E * Q, D), W.p(pre + "norm3.weight"), W.p(pre + "norm3.bias"))
    cache = None
    if need_cache:
        cache = dict(qk3=qk3, v3=v3, P=P, o=o, a1=a1, m1=m1, r1=r1, qc=qc, kc=kc, vc=vc, P2=P2, o2=o2,
                     a2=a2, m2=m2, r2=r2, t2_r=t2_r, h=h, f=f, m3=m3, r3=r3,
                     qk_in=qk_in, tgt_r=tgt_r, q_in=q_in, mem_pos_r=mem_pos_r, memory_r=memory_r,
                     drop=None if drop is None else (kd1, kd2, kf, kd3))
    if want_plus:
        return t3.view(E, Q, D), t3_r.view(E, Q, D), cache, t3_plus.view(E, Q, D)
    return t3.view(E, Q, D), t3_r.view(E, Q, D), cache


def decoder_layer_bwd(ops, W, pre, dm, s, dt, sink, dqpos, dmp, dmem):
    """dt [E*Q,D] = gradient of the layer output.  Returns the gradient of the layer input
    [E*Q,D]; accumulates d(query_pos) into dqpos [E,Q,D] (if given), d(memory+pos) into dmp and
    d(memory) into dmem (both [1,E*R,D]); writes weight gradients through `sink` (NullSink = data
    gradients only, as for the fusion network on the inner loop, whose parameters are not adapted)."""
    E, B, Lq, Lk, D, nh, hd, Q, R = dm.E, dm.B, dm.Lq, dm.Lk, dm.D, dm.nh, dm.hd, dm.Q, dm.R
    sa_w, ca_w = pre + "self_attn.in_proj_weight", pre + "multihead_attn.in_proj_weight"
    sa_b, ca_b = pre + "self_attn.in_proj_bias", pre + "multihead_attn.in_proj_bias"
    sink = sink if sink is not None else NullSink()
    dk_ = s.get("drop")
    # without dropout the residual branch's gradient IS the LayerNorm's dx: its linear's bias gradient is fused
    nk = lambda n, lin: sink.norm(pre + n, bias_of=pre + lin if dk_ is None else None)
    df, df_r = ops.layernorm_bwd(dt, s["f"].view(E * Q, D), s["m3"], s["r3"], W.p(pre + "norm3.weight"),
                                 **nk("norm3", "linear2"))
    df3, df3_r = df.view(E, Q, D), df_r.view(E, Q, D)
    kd1, kd2, kf, kd3 = dk_ if dk_ is not None else (None,) * 4
    if dk_ is not None:                       # gradient of the dropped branch; the residual path keeps df3
        dfm = ops.dropout(df3, kd3)
        dfm_r = dfm
    else:
        dfm, dfm_r = df3, df3_r
    dh = ops.matmul(dfm_r, W.bwd(pre + "linear2.weight"), epi="relu_mask", aux=s["h"], rnd=True)
    if dk_ is not None:
        dh = ops.dropout(dh, kf, out=dh)      # h is stored dropped: relu_mask kept (h > 0) & keep, this adds 1/(1-p)
    sink.linear(pre + "linear2", dfm_r, s["h"], dfm)
    sink.linear(pre + "linear1", dh, s["t2_r"].view(E, Q, D))
    dt2 = ops.matmul(dh, W.bwd(pre + "linear1.weight"), residual=df3)
    da2, da2_r = ops.layernorm_bwd(dt2.view(E * Q, D), s["a2"].view(E * Q, D), s["m2"], s["r2"],
                                   W.p(pre + "norm2.weight"), **nk("norm2", "multihead_attn.out_proj"))
    da2_3r = da2_r.view(E, Q, D)
    da2m = da2.view(E, Q, D)
    if dk_ is not None:
        da2m = ops.dropout(da2m, kd2)
        da2_3r = da2m
    sink.linear(pre + "multihead_attn.out_proj", da2_3r, s["o2"].view(E, Q, D), da2m)
    dO2 = ops.matmul(da2_3r, W.bwd(pre + "multihead_attn.out_proj.weight"), rnd=True)
    dqc, dkc, dvc = ops.empty(B, Lq, D), ops.empty(B, Lk, D), ops.empty(B, Lk, D)
    attention_bwd(ops, dO2.view(B, Lq, D), s["qc"], s["kc"], s["vc"], s["P2"], B, Lq, Lk, nh, hd, dm.scale,
                  dqc, dkc, dvc)
    dqc1 = dqc.view(1, E * Q, D)
    if sink.wants(ca_w):
        sink.rows(ca_w, ca_b, 0, D, dqc1, s["q_in"])
        sink.rows(ca_w, ca_b, D, 2 * D, dkc.view(1, E * R, D), s["mem_pos_r"])
        sink.rows(ca_w, ca_b, 2 * D, 3 * D, dvc.view(1, E * R, D), s["memory_r"])
    dt1 = ops.matmul(dqc1, W.bwd(ca_w, 0, D), residual=da2.view(1, E * Q, D))
    if dqpos is not None:
        ops.matmul(dqc1, W.bwd(ca_w, 0, D), out=dqpos.view(1, E * Q, D), accumulate=True)
    ops.matmul(dkc.view(1, E * R, D), W.bwd(ca_w, D, 2 * D), out=dmp, accumulate=True)
    ops.matmul(dvc.view(1, E * R, D), W.bwd(ca_w, 2 * D, 3 * D), out=dmem, accumulate=True)
    da1, da1_r = ops.layernorm_bwd(dt1.view(E * Q, D), s["a1"].view(E * Q, D), s["m1"], s["r1"],
                                   W.p(pre + "norm1.weight"), **nk("norm1", "self_attn.out_proj"))
    da1_3r = da1_r.view(E, Q, D)
    da1m = da1.view(E, Q, D)
    if dk_ is not None:
        da1m = ops.dropout(da1m, kd1)
        da1_3r = da1m
    sink.linear(pre + "self_attn.out_proj", da1_3r, s["o"].view(E, Q, D), da1m)
    dO = ops.matmul(da1_3r, W.bwd(pre + "self_attn.out_proj.weight"), rnd=True)
    dqk, dv = ops.empty(B, Lq, 2 * D), ops.empty(B, Lq, D)
    attention_bwd(ops, dO.view(B, Lq, D), s["qk3"][..., :D], s["qk3"][..., D:], s["v3"], s["P"],
                  B, Lq, Lq, nh, hd, dm.scale, dqk[..., :D], dqk[..., D:], dv)
    dqk1 = dqk.view(1, E * Q, 2 * D)
    if sink.wants(sa_w):
        sink.rows(sa_w, sa_b, 0, 2 * D, dqk1, s["qk_in"])
        sink.rows(sa_w, sa_b, 2 * D, 3 * D, dv.view(1, E * Q, D), s["tgt_r"].view(1, E * Q, D))
    dtg = ops.matmul(dqk1, W.bwd(sa_w, 0, 2 * D), residual=da1.view(1, E * Q, D))
    if dqpos is not None:
        ops.matmul(dqk1, W.bwd(sa_w, 0, 2 * D), out=dqpos.view(1, E * Q, D), accumulate=True)
    ops.matmul(dv.view(1, E * Q, D), W.bwd(sa_w, 2 * D, 3 * D), out=dtg, accumulate=True)
    return dtg.view(E * Q, D)


# --------------------------------------------------------------------------- MLP heads
def mlp_fwd(ops, W, name, x_r, n_layers=3):
    """DETR-style MLP (reference detr_models/detr.py:299-311): ReLU between layers, none after the
    last.  x_r [E,R,K] TF32-clean -> (z [E,R,out] full fp32, hidden activations for the backward)."""
    hid = []
    h = x_r
    for i in range(n_layers - 1):
        h = lin(ops, h, W.w(f"{name}.layers.{i}.weight"), W.p(f"{name}.layers.{i}.bias"), act="relu", rnd=True)
        hid.append(h)
    z = lin(ops, h, W.w(f"{name}.layers.{n_layers - 1}.weight"), W.p(f"{name}.layers.{n_layers - 1}.bias"))
    return z, hid


def mlp_bwd(ops, W, name, dz_r, x_r, hid, sink=None, n_layers=3, **last_kw):
    """dz_r [E,R,out] TF32-clean -> gradient wrt x (extra epilogue args via last_kw)."""
    d = dz_r
    sink = sink if sink is not None else NullSink()
    for i in reversed(range(n_layers)):
        inp = hid[i - 1] if i > 0 else x_r
        sink.linear(f"{name}.layers.{i}", d, inp)
        w = W.bwd(f"{name}.layers.{i}.weight")
        if i > 0:
            d = ops.matmul(d, w, epi="relu_mask", aux=hid[i - 1], rnd=True)
        else:
            d = ops.matmul(d, w, **last_kw)
    return d
