"""DETR transformer + heads (input_proj -> 6 encoder -> 6 decoder layers -> class/box heads):
forward and hand-derived backward, written against the `ops` kernel interface.

Replaces, on the hot path, reference models/detr_models/detr.py:66-75 (input_proj, heads),
models/detr_models/transformer.py:46-58,148-161,211-232 (post-norm encoder / decoder layers)
and the autograd graph that models/interactron.py:51-52 differentiates.

Layout: activations are token-major [E, rows, D] (E episodes, rows = frames*tokens), never the
reference's [tokens, batch, D]; heads are strided views, no transposes are materialised.
Weights come from `Weights` views with a leading group dim (1 = shared theta, E = per-episode
fast weights theta').  in_proj_* of every nn.MultiheadAttention is not a fast weight
(reference utils/meta_utils.py:9-21) and is always shared.

TF32 hygiene: tensors that feed a GEMM are stored TF32-rounded by their producer (`rnd=True`,
the `_r` twins of LayerNorm outputs); residual streams and gradients accumulate in full fp32.
"""
import math

D, H, HD, FFN, NQ = 256, 8, 32, 2048, 50
N_ENC, N_DEC = 6, 6
SCALE = 1.0 / math.sqrt(HD)


def pad4(n):
    return (n + 3) // 4 * 4


def T(w):
    return w.transpose(-1, -2)


# --------------------------------------------------------------------------- attention core
def attention_fwd(ops, q, k, v, B, Lq, Lk, nh, hd, scale, kmask):
    """q [B,Lq,nh*hd], k/v [B,Lk,nh*hd] (strided views ok) -> o [B,Lq,nh*hd] (TF32-clean), P."""
    qh = q.reshape(B, Lq, nh, hd).permute(0, 2, 1, 3)            # [B,nh,Lq,hd]
    khT = k.reshape(B, Lk, nh, hd).permute(0, 2, 3, 1)           # [B,nh,hd,Lk]
    vh = v.reshape(B, Lk, nh, hd).permute(0, 2, 1, 3)            # [B,nh,Lk,hd]
    P = ops.empty(B, nh, Lq, pad4(Lk))                           # row stride padded to 16 bytes for TMA
    p = P[..., :Lk]
    ops.matmul(qh, khT, out=p)
    ops.softmax_(P, Lk, scale, kmask, rows_per_mask=nh * Lq)
    o = ops.empty(B, Lq, nh * hd)
    ops.matmul(p, vh, out=o.view(B, Lq, nh, hd).permute(0, 2, 1, 3), rnd=True)
    return o, P


def attention_bwd(ops, dO, q, k, v, P, B, Lq, Lk, nh, hd, scale, dq, dk, dv):
    """dO [B,Lq,nh*hd] (TF32-clean).  Writes TF32-clean dq/dk/dv into the given [B,L,nh*hd] views."""
    qh = q.reshape(B, Lq, nh, hd).permute(0, 2, 1, 3)
    kh = k.reshape(B, Lk, nh, hd).permute(0, 2, 1, 3)
    vhT = v.reshape(B, Lk, nh, hd).permute(0, 2, 3, 1)
    dOh = dO.reshape(B, Lq, nh, hd).permute(0, 2, 1, 3)
    p = P[..., :Lk]
    dP = ops.empty(B, nh, Lq, pad4(Lk))
    dp = dP[..., :Lk]
    ops.matmul(dOh, vhT, out=dp)                                                    # dP = dO V^T
    ops.matmul(T(p), dOh, out=dv.view(B, Lk, nh, hd).permute(0, 2, 1, 3), rnd=True)  # dV = P^T dO
    ops.softmax_bwd_(P, dP, Lk, scale)                                              # dS (in dP)
    ops.matmul(dp, kh, out=dq.view(B, Lq, nh, hd).permute(0, 2, 1, 3), rnd=True)     # dQ = dS K
    ops.matmul(T(dp), qh, out=dk.view(B, Lk, nh, hd).permute(0, 2, 1, 3), rnd=True)  # dK = dS^T Q


# --------------------------------------------------------------------------- helpers
def lin(ops, x, W, b=None, **kw):
    """x [E,R,K] @ W[Gw,N,K]^T (+ b[Gw,N]) -> [E,R,N]."""
    return ops.matmul(x, T(W), bias=b, **kw)


class GradSink:
    """Writes weight gradients straight into the flat per-episode gradient buffer g [E, n_theta]."""

    def __init__(self, ops, pack, g):
        self.ops, self.pack, self.g = ops, pack, g

    def view(self, name):
        return self.pack.view(self.g, name)

    def linear(self, name, dy_r, x_r, dy_full=None):
        """dW = dy^T x into `name.weight`, db = colsum(dy) into `name.bias`."""
        w = self.view(name + ".weight")
        E = w.shape[0]
        self.ops.matmul(T(dy_r), x_r, out=w.reshape(E, w.shape[1], -1))
        self.ops.colsum(dy_full if dy_full is not None else dy_r, out=self.view(name + ".bias"))


# --------------------------------------------------------------------------- forward
def detr_t_forward(ops, W, src_r, pos, kmask, E, Fe, L, preds=None, need_cache=True):
    """
    src_r : [E, Fe*L, 2048] TF32-clean backbone features, token-major
    pos   : [E*Fe*L, 256] sine position embedding (full fp32)
    kmask : uint8 [E*Fe, L] (1 = padded key) or None
    preds : optional [E*Fe*50, 1496] buffer that receives cat(box_features, logits, boxes)
            TF32-clean, i.e. the prediction tokens fusion embeds (reference models/transformer.py:50)
    Returns (out, cache): out has logits [E,Fe*50,C], boxes [E,Fe*50,4], hs, memory (+ `_r` twins).
    """
    R, Q, B = Fe * L, Fe * NQ, E * Fe
    c = {} if need_cache else None
    ip_w = W.w("input_proj.weight")
    ip_w = ip_w.reshape(ip_w.shape[0], D, -1)
    x = ops.empty(E, R, D)
    x_r = lin(ops, src_r, ip_w, W.p("input_proj.bias"), out_pre=x, rnd=True)        # x (fp32), x_r
    enc = []
    for i in range(N_ENC):
        pre = f"transformer.encoder.layers.{i}."
        ipw, ipb = W.w(pre + "self_attn.in_proj_weight"), W.p(pre + "self_attn.in_proj_bias")
        qk_in = ops.add(x.view(E * R, D), pos, rnd=True).view(1, E * R, D)
        qk = lin(ops, qk_in, ipw[:, :2 * D], ipb[:, :2 * D], rnd=True)               # [1,E*R,512]
        v = lin(ops, x_r.view(1, E * R, D), ipw[:, 2 * D:], ipb[:, 2 * D:], rnd=True)
        qk3, v3 = qk.view(B, L, 2 * D), v.view(B, L, D)
        o, P = attention_fwd(ops, qk3[..., :D], qk3[..., D:], v3, B, L, L, H, HD, SCALE, kmask)
        a = lin(ops, o.view(E, R, D), W.w(pre + "self_attn.out_proj.weight"),
                W.p(pre + "self_attn.out_proj.bias"), residual=x)
        x1, x1_r, m1, r1 = ops.layernorm_fwd(a.view(E * R, D), W.p(pre + "norm1.weight"),
                                             W.p(pre + "norm1.bias"))
        h = lin(ops, x1_r.view(E, R, D), W.w(pre + "linear1.weight"), W.p(pre + "linear1.bias"),
                act="relu", rnd=True)
        f = lin(ops, h, W.w(pre + "linear2.weight"), W.p(pre + "linear2.bias"), residual=x1.view(E, R, D))
        x2, x2_r, m2, r2 = ops.layernorm_fwd(f.view(E * R, D), W.p(pre + "norm2.weight"),
                                             W.p(pre + "norm2.bias"))
        if need_cache:
            enc.append(dict(x_r=x_r, qk_in=qk_in, qk3=qk3, v3=v3, P=P, o=o, a=a, m1=m1, r1=r1,
                            x1_r=x1_r, h=h, f=f, m2=m2, r2=r2))
        x, x_r = x2.view(E, R, D), x2_r.view(E, R, D)
    memory, memory_r = x, x_r
    mem_pos_r = ops.add(memory.view(E * R, D), pos, rnd=True).view(1, E * R, D)

    qpos = W.p("query_embed.weight")                                                  # [Gw,50,256]
    tgt = ops.zeros(E, Q, D)
    tgt_r = tgt
    dec = []
    for j in range(N_DEC):
        pre = f"transformer.decoder.layers.{j}."
        sw, sb = W.w(pre + "self_attn.in_proj_weight"), W.p(pre + "self_attn.in_proj_bias")
        cw, cb = W.w(pre + "multihead_attn.in_proj_weight"), W.p(pre + "multihead_attn.in_proj_bias")
        # self attention over the 50 queries of each frame
        qk_in = ops.add(tgt, qpos, rnd=True).view(1, E * Q, D)
        qk = lin(ops, qk_in, sw[:, :2 * D], sb[:, :2 * D], rnd=True)
        v = lin(ops, tgt_r.view(1, E * Q, D), sw[:, 2 * D:], sb[:, 2 * D:], rnd=True)
        qk3, v3 = qk.view(B, NQ, 2 * D), v.view(B, NQ, D)
        o, P = attention_fwd(ops, qk3[..., :D], qk3[..., D:], v3, B, NQ, NQ, H, HD, SCALE, None)
        a1 = lin(ops, o.view(E, Q, D), W.w(pre + "self_attn.out_proj.weight"),
                 W.p(pre + "self_attn.out_proj.bias"), residual=tgt)
        t1, t1_r, m1, r1 = ops.layernorm_fwd(a1.view(E * Q, D), W.p(pre + "norm1.weight"),
                                             W.p(pre + "norm1.bias"))
        # cross attention into the frame's 361 memory tokens
        q_in = ops.add(t1.view(E, Q, D), qpos, rnd=True).view(1, E * Q, D)
        qc = lin(ops, q_in, cw[:, :D], cb[:, :D], rnd=True).view(B, NQ, D)
        kc = lin(ops, mem_pos_r, cw[:, D:2 * D], cb[:, D:2 * D], rnd=True).view(B, L, D)
        vc = lin(ops, memory_r.view(1, E * R, D), cw[:, 2 * D:], cb[:, 2 * D:], rnd=True).view(B, L, D)
        o2, P2 = attention_fwd(ops, qc, kc, vc, B, NQ, L, H, HD, SCALE, kmask)
        a2 = lin(ops, o2.view(E, Q, D), W.w(pre + "multihead_attn.out_proj.weight"),
                 W.p(pre + "multihead_attn.out_proj.bias"), residual=t1.view(E, Q, D))
        t2, t2_r, m2, r2 = ops.layernorm_fwd(a2.view(E * Q, D), W.p(pre + "norm2.weight"),
                                             W.p(pre + "norm2.bias"))
        h = lin(ops, t2_r.view(E, Q, D), W.w(pre + "linear1.weight"), W.p(pre + "linear1.bias"),
                act="relu", rnd=True)
        f = lin(ops, h, W.w(pre + "linear2.weight"), W.p(pre + "linear2.bias"), residual=t2.view(E, Q, D))
        t3, t3_r, m3, r3 = ops.layernorm_fwd(f.view(E * Q, D), W.p(pre + "norm3.weight"),
                                             W.p(pre + "norm3.bias"))
        if need_cache:
            dec.append(dict(tgt_r=tgt_r, qk_in=qk_in, qk3=qk3, v3=v3, P=P, o=o, a1=a1, m1=m1, r1=r1,
                            t1_r=t1_r, q_in=q_in, qc=qc, kc=kc, vc=vc, P2=P2, o2=o2, a2=a2, m2=m2, r2=r2,
                            t2_r=t2_r, h=h, f=f, m3=m3, r3=r3))
        tgt, tgt_r = t3.view(E, Q, D), t3_r.view(E, Q, D)

    # only hs[-1] is used (reference detr.py:69), so decoder.norm runs once
    hs, hs_r, mh, rh = ops.layernorm_fwd(tgt.view(E * Q, D), W.p("transformer.decoder.norm.weight"),
                                         W.p("transformer.decoder.norm.bias"))
    hs, hs_r = hs.view(E, Q, D), hs_r.view(E, Q, D)
    cw_, cb_ = W.w("class_embed.weight"), W.p("class_embed.bias")
    C = cw_.shape[1]
    logits = ops.empty(E, Q, C)
    if preds is not None:
        pv = preds.view(E, Q, D + C + 4)
        lin(ops, hs_r, cw_, cb_, out=pv[..., D:D + C], out_pre=logits, rnd=True)
        ops.copy2d_(preds[:, :D], hs_r.view(E * Q, D))
    else:
        lin(ops, hs_r, cw_, cb_, out=logits)
    b1 = lin(ops, hs_r, W.w("bbox_embed.layers.0.weight"), W.p("bbox_embed.layers.0.bias"), act="relu", rnd=True)
    b2 = lin(ops, b1, W.w("bbox_embed.layers.1.weight"), W.p("bbox_embed.layers.1.bias"), act="relu", rnd=True)
    z = lin(ops, b2, W.w("bbox_embed.layers.2.weight"), W.p("bbox_embed.layers.2.bias"))
    boxes = ops.sigmoid(z)
    if preds is not None:
        ops.copy2d_(preds[:, D + C:], boxes.view(E * Q, 4), rnd=True)
    out = dict(logits=logits, boxes=boxes, hs=hs, hs_r=hs_r, memory=memory, memory_r=memory_r)
    if need_cache:
        c.update(enc=enc, dec=dec, src_r=src_r, tgt_last=tgt, mh=mh, rh=rh, hs_r=hs_r, b1=b1, b2=b2,
                 boxes=boxes, mem_pos_r=mem_pos_r, memory_r=memory_r, E=E, Fe=Fe, L=L)
    return out, c


# --------------------------------------------------------------------------- backward
def detr_t_backward(ops, W, cache, sink, dpreds=None, dmemory=None, dlogits=None, dboxes=None, dhs=None):
    """Back-propagates into the flat gradient buffer behind `sink` (all 157 fast weights).

    Upstream gradients: either `dpreds` [E*Fe*50, 1496] (TF32-clean gradient of the fusion
    prediction tokens = cat(d box_features, d logits, d boxes)) and `dmemory` [E, Fe*L, 256]
    (gradient of the encoder memory from fusion), or explicit dlogits/dboxes/dhs.
    in_proj_* receive no gradient here (they are not fast weights); the backbone is frozen (D1).
    """
    E, Fe, L = cache["E"], cache["Fe"], cache["L"]
    R, Q, B = Fe * L, Fe * NQ, E * Fe
    C = W.w("class_embed.weight").shape[1]
    if dpreds is not None:
        pv = dpreds.view(E, Q, D + C + 4)
        dhs_in, dlogits_r, dboxes = pv[..., :D], pv[..., D:D + C], pv[..., D + C:]
    else:
        dhs_in, dlogits_r = dhs, dlogits
    hs_r = cache["hs_r"]

    # heads ---------------------------------------------------------------------------
    dbx = ops.copy2d_(ops.empty(E * Q, 4), dboxes.reshape(E * Q, 4)).view(E, Q, 4)
    dz = ops.sigmoid_bwd(dbx, cache["boxes"])                                           # [E,Q,4]
    dz = ops.round_tf32(dz, out=dz)
    sink.linear("bbox_embed.layers.2", dz, cache["b2"])
    db2 = ops.matmul(dz, W.w("bbox_embed.layers.2.weight"), epi="relu_mask", aux=cache["b2"], rnd=True)
    sink.linear("bbox_embed.layers.1", db2, cache["b1"])
    db1 = ops.matmul(db2, W.w("bbox_embed.layers.1.weight"), epi="relu_mask", aux=cache["b1"], rnd=True)
    sink.linear("bbox_embed.layers.0", db1, hs_r)
    dhs_t = ops.matmul(db1, W.w("bbox_embed.layers.0.weight"), residual=dhs_in)
    sink.linear("class_embed", dlogits_r, hs_r)
    ops.matmul(dlogits_r, W.w("class_embed.weight"), out=dhs_t, accumulate=True)
    dt, _ = ops.layernorm_bwd(dhs_t.view(E * Q, D), cache["tgt_last"].view(E * Q, D), cache["mh"], cache["rh"],
                              W.p("transformer.decoder.norm.weight"),
                              dgamma=sink.view("transformer.decoder.norm.weight"),
                              dbeta=sink.view("transformer.decoder.norm.bias"))

    # decoder -------------------------------------------------------------------------
    dmem = dmemory.contiguous() if dmemory is not None else ops.zeros(E, R, D)         # accumulates
    dmp = ops.zeros(1, E * R, D)                                                        # d(memory+pos)
    dqpos = ops.zeros(E, Q, D)                                                          # d(query_pos), per frame
    mem_pos_r, memory_r = cache["mem_pos_r"], cache["memory_r"]
    for j in reversed(range(N_DEC)):
        pre = f"transformer.decoder.layers.{j}."
        s = cache["dec"][j]
        sw = W.w(pre + "self_attn.in_proj_weight")
        cw = W.w(pre + "multihead_attn.in_proj_weight")
        df, df_r = ops.layernorm_bwd(dt, s["f"].view(E * Q, D), s["m3"], s["r3"], W.p(pre + "norm3.weight"),
                                     dgamma=sink.view(pre + "norm3.weight"), dbeta=sink.view(pre + "norm3.bias"))
        df3, df3_r = df.view(E, Q, D), df_r.view(E, Q, D)
        sink.linear(pre + "linear2", df3_r, s["h"], df3)
        dh = ops.matmul(df3_r, W.w(pre + "linear2.weight"), epi="relu_mask", aux=s["h"], rnd=True)
        sink.linear(pre + "linear1", dh, s["t2_r"].view(E, Q, D))
        dt2 = ops.matmul(dh, W.w(pre + "linear1.weight"), residual=df3)
        da2, da2_r = ops.layernorm_bwd(dt2.view(E * Q, D), s["a2"].view(E * Q, D), s["m2"], s["r2"],
                                       W.p(pre + "norm2.weight"), dgamma=sink.view(pre + "norm2.weight"),
                                       dbeta=sink.view(pre + "norm2.bias"))
        da2_3, da2_3r = da2.view(E, Q, D), da2_r.view(E, Q, D)
        sink.linear(pre + "multihead_attn.out_proj", da2_3r, s["o2"].view(E, Q, D), da2_3)
        dO2 = ops.matmul(da2_3r, W.w(pre + "multihead_attn.out_proj.weight"), rnd=True)
        dqc, dkc, dvc = ops.empty(B, NQ, D), ops.empty(B, L, D), ops.empty(B, L, D)
        attention_bwd(ops, dO2.view(B, NQ, D), s["qc"], s["kc"], s["vc"], s["P2"], B, NQ, L, H, HD, SCALE,
                      dqc, dkc, dvc)
        dqc1 = dqc.view(1, E * Q, D)
        dt1 = ops.matmul(dqc1, cw[:, :D], residual=da2.view(1, E * Q, D))               # d t1
        ops.matmul(dqc1, cw[:, :D], out=dqpos.view(1, E * Q, D), accumulate=True)
        ops.matmul(dkc.view(1, E * R, D), cw[:, D:2 * D], out=dmp, accumulate=True)
        ops.matmul(dvc.view(1, E * R, D), cw[:, 2 * D:], out=dmem.view(1, E * R, D), accumulate=True)
        da1, da1_r = ops.layernorm_bwd(dt1.view(E * Q, D), s["a1"].view(E * Q, D), s["m1"], s["r1"],
                                       W.p(pre + "norm1.weight"), dgamma=sink.view(pre + "norm1.weight"),
                                       dbeta=sink.view(pre + "norm1.bias"))
        da1_3, da1_3r = da1.view(E, Q, D), da1_r.view(E, Q, D)
        sink.linear(pre + "self_attn.out_proj", da1_3r, s["o"].view(E, Q, D), da1_3)
        dO = ops.matmul(da1_3r, W.w(pre + "self_attn.out_proj.weight"), rnd=True)
        dqk, dv = ops.empty(B, NQ, 2 * D), ops.empty(B, NQ, D)
        attention_bwd(ops, dO.view(B, NQ, D), s["qk3"][..., :D], s["qk3"][..., D:], s["v3"], s["P"],
                      B, NQ, NQ, H, HD, SCALE, dqk[..., :D], dqk[..., D:], dv)
        dqk1 = dqk.view(1, E * Q, 2 * D)
        dtg = ops.matmul(dqk1, sw[:, :2 * D], residual=da1.view(1, E * Q, D))
        ops.matmul(dqk1, sw[:, :2 * D], out=dqpos.view(1, E * Q, D), accumulate=True)
        ops.matmul(dv.view(1, E * Q, D), sw[:, 2 * D:], out=dtg, accumulate=True)
        dt = dtg.view(E * Q, D)
    # query_embed gradient: sum the per-frame query_pos gradients of each episode
    ops.colsum(dqpos.view(E, Fe, NQ * D), out=sink.view("query_embed.weight").reshape(E, NQ * D))
    # gradient of memory = decoder V paths (dmem) + K paths through memory+pos (dmp) [+ fusion]
    dx = ops.add(dmem.view(E * R, D), dmp.view(E * R, D))

    # encoder -------------------------------------------------------------------------
    for i in reversed(range(N_ENC)):
        pre = f"transformer.encoder.layers.{i}."
        s = cache["enc"][i]
        ipw = W.w(pre + "self_attn.in_proj_weight")
        df, df_r = ops.layernorm_bwd(dx, s["f"].view(E * R, D), s["m2"], s["r2"], W.p(pre + "norm2.weight"),
                                     dgamma=sink.view(pre + "norm2.weight"), dbeta=sink.view(pre + "norm2.bias"))
        df3, df3_r = df.view(E, R, D), df_r.view(E, R, D)
        sink.linear(pre + "linear2", df3_r, s["h"], df3)
        dh = ops.matmul(df3_r, W.w(pre + "linear2.weight"), epi="relu_mask", aux=s["h"], rnd=True)
        sink.linear(pre + "linear1", dh, s["x1_r"].view(E, R, D))
        dx1 = ops.matmul(dh, W.w(pre + "linear1.weight"), residual=df3)
        da, da_r = ops.layernorm_bwd(dx1.view(E * R, D), s["a"].view(E * R, D), s["m1"], s["r1"],
                                     W.p(pre + "norm1.weight"), dgamma=sink.view(pre + "norm1.weight"),
                                     dbeta=sink.view(pre + "norm1.bias"))
        da3, da3_r = da.view(E, R, D), da_r.view(E, R, D)
        sink.linear(pre + "self_attn.out_proj", da3_r, s["o"].view(E, R, D), da3)
        dO = ops.matmul(da3_r, W.w(pre + "self_attn.out_proj.weight"), rnd=True)
        dqk, dv = ops.empty(B, L, 2 * D), ops.empty(B, L, D)
        attention_bwd(ops, dO.view(B, L, D), s["qk3"][..., :D], s["qk3"][..., D:], s["v3"], s["P"],
                      B, L, L, H, HD, SCALE, dqk[..., :D], dqk[..., D:], dv)
        dxi = ops.matmul(dqk.view(1, E * R, 2 * D), ipw[:, :2 * D], residual=da.view(1, E * R, D))
        last = i == 0
        ops.matmul(dv.view(1, E * R, D), ipw[:, 2 * D:], out=dxi, accumulate=True, rnd=last)
        dx = dxi.view(E * R, D)

    # input_proj ----------------------------------------------------------------------
    dx3 = dx.view(E, R, D)                                            # TF32-clean (rounded by the last GEMM)
    wv = sink.view("input_proj.weight")
    ops.matmul(T(dx3), cache["src_r"], out=wv.reshape(E, D, -1))
    ops.colsum(dx3, out=sink.view("input_proj.bias"))
