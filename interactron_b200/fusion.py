"""The two "fusion" networks that turn a 5-frame episode of detector outputs into the learned
loss: forward and hand-derived data-gradient backward over the `ops` kernel interface.

  fusion A (`interactron`)        reference models/transformer.py:47-66 + models/gpt.py:39-78,189-200
      2060-token sequence [1805 image tokens | 250 prediction tokens | 5 action tokens] through a
      4-layer pre-LN GPT (d=512, 8 heads, full attention, exact-erf GELU).
  fusion B (`interactron_random`) reference models/new_transformer.py:34-58
      255 target tokens (250 predictions + 5 action tokens) attend 1805 image tokens through
      4 DETR decoder layers (d=512, 8 heads, ReLU FFN 2048).

Inputs come from detr_t.detr_t_forward: `memory_r` [E,S*361,256] and the TF32-clean prediction
tokens `preds` [E*S*50, 1496] = cat(box_features, logits, boxes).  Only *data* gradients are
needed on the inner loop (the fusion parameters phi are not adapted): `backward` returns
d memory and d preds for detr_t.detr_t_backward.
"""
from .layers import (DecDims, NullSink, T, attention_bwd, attention_fwd, decoder_layer_bwd,  # noqa: F401
                     cross_kv_all, decoder_layer_fwd, lin, mlp_bwd, mlp_fwd, _drop_res, _next)

DF, NH, NP, NA = 512, 8, 50, 5          # width, heads, predictions per frame, action tokens
N_LAYERS = 4


def _learned_loss(ops, W, y_r, E, S, need_cache):
    """loss_decoder MLP on the S*50 prediction tokens + L2 norm (reference models/interactron.py:50).
    y_r [E, S*50, 512] TF32-clean -> (loss_vec [E,S*50], learned_loss [E], dloss_vec, hidden)."""
    z, hid = mlp_fwd(ops, W, "loss_decoder", y_r)
    lv = z.view(E, S * NP)
    norm, dz = ops.l2norm_fwd_bwd(lv)
    return lv, norm, dz, hid


# ------------------------------------------------------------------------------------- fusion B
def fusion_b_forward(ops, W, memory_r, preds, E, S, L, need_cache=True, drop=None):
    """-> dict(loss_vec [E,S*50], learned_loss [E], actions [E,4,4]), cache."""
    assert S == 5, "fusion B is only defined for full 5-frame episodes"
    R, Qp, Q = S * L, S * NP, S * NP + NA
    mem = lin(ops, memory_r.view(1, E * R, -1), W.w("img_feature_embedding.weight"),
              W.p("img_feature_embedding.bias"))                                     # [1,E*R,512] fp32
    mem_r = ops.round_tf32(mem)
    mem_pos_r = ops.add(mem.view(E, R * DF), W.p("pos_embed").reshape(1, R * DF), rnd=True).view(1, E * R, DF)
    tgt = ops.empty(E, Q, DF)
    lin(ops, preds.view(E, Qp, -1), W.w("prediction_embedding.weight"), W.p("prediction_embedding.bias"),
        out=tgt[:, :Qp])
    at = W.p("action_tokens")                                                          # [1,1,5,512]
    ops.copy2d_(tgt.view(E, Q * DF)[:, Qp * DF:], at.reshape(1, NA * DF).expand(E, NA * DF))
    tgt_r = ops.round_tf32(tgt)
    qpos = W.p("query_embed").reshape(1, Q, DF)
    dm = DecDims(E, E, Q, R, DF, NH)
    caches = []
    x, x_r = tgt, tgt_r
    xq = None                        # x + query_pos, fused into the previous layer's norm3
    kvs = cross_kv_all(ops, W, [f"transformer.layers.{j}." for j in range(N_LAYERS)], dm, mem_pos_r, mem_r)
    for j in range(N_LAYERS):
        res = decoder_layer_fwd(ops, W, f"transformer.layers.{j}.", dm, x, x_r, qpos, mem_pos_r, mem_r,
                                None, need_cache, drop=drop, tgt_plus=xq, want_plus=j < N_LAYERS - 1,
                                kv=None if kvs is None else kvs[j])
        x, x_r, c = res[:3]
        xq = res[3] if j < N_LAYERS - 1 else None
        caches.append(c)
    y, y_r, my, ry = ops.layernorm_fwd(x.view(E * Q, DF), W.p("transformer.norm.weight"),
                                       W.p("transformer.norm.bias"))
    y_r = y_r.view(E, Q, DF)
    yp_r = ops.copy2d_(ops.empty(E, Qp * DF), y_r.view(E, Q * DF)[:, :Qp * DF]).view(E, Qp, DF)
    ya_r = ops.copy2d_(ops.empty(E, 4 * DF), y_r.view(E, Q * DF)[:, Qp * DF:(Qp + 4) * DF]).view(E, 4, DF)
    lv, norm, dlv, lhid = _learned_loss(ops, W, yp_r, E, S, need_cache)
    actions, ahid = mlp_fwd(ops, W, "action_decoder", ya_r)
    out = dict(loss_vec=lv, learned_loss=norm, actions=actions)
    cache = None
    if need_cache:
        cache = dict(layers=caches, x_last=x, my=my, ry=ry, yp_r=yp_r, dlv=dlv, lhid=lhid, E=E, S=S, L=L,
                     ya_r=ya_r, ahid=ahid, memory_r=memory_r, preds=preds)
    return out, cache


def fusion_b_backward(ops, W, cache, sink=None, dactions=None):
    """Gradient of sum_e learned_loss[e] (+ <dactions, actions> when given) wrt the inputs:
    -> (dmemory [E,S*L,256] fp32, dpreds [E*S*50,1496] TF32-clean).  `sink` (default: none, the
    inner loop) receives the gradients of the fusion parameters phi (meta-training step)."""
    E, S, L = cache["E"], cache["S"], cache["L"]
    R, Qp, Q = S * L, S * NP, S * NP + NA
    sink = sink if sink is not None else NullSink()
    dz = ops.round_tf32(cache["dlv"].view(E, Qp, 1))
    dyp = mlp_bwd(ops, W, "loss_decoder", dz, cache["yp_r"], cache["lhid"], sink)      # [E,Qp,512]
    dy = ops.zeros(E, Q, DF)
    ops.copy2d_(dy.view(E, Q * DF)[:, :Qp * DF], dyp.view(E, Qp * DF))
    if dactions is not None:
        dya = mlp_bwd(ops, W, "action_decoder", dactions, cache["ya_r"], cache["ahid"], sink)   # [E,4,512]
        ops.copy2d_(dy.view(E, Q * DF)[:, Qp * DF:(Qp + 4) * DF], dya.view(E, 4 * DF))
    dt, _ = ops.layernorm_bwd(dy.view(E * Q, DF), cache["x_last"].view(E * Q, DF), cache["my"], cache["ry"],
                              W.p("transformer.norm.weight"), **sink.norm("transformer.norm"))
    dm = DecDims(E, E, Q, R, DF, NH)
    dmp, dmem = ops.zeros(1, E * R, DF), ops.zeros(1, E * R, DF)
    dqpos = ops.zeros(E, Q, DF) if sink.wants("query_embed") else None
    for j in reversed(range(N_LAYERS)):
        dt = decoder_layer_bwd(ops, W, f"transformer.layers.{j}.", dm, cache["layers"][j], dt, sink, dqpos,
                               dmp, dmem)
    if dqpos is not None:
        sink.colsum("query_embed", dqpos.view(E, 1, Q * DF))
    sink.colsum("action_tokens", dt.view(1, E, Q * DF)[:, :, Qp * DF:])
    dmem_tot = ops.add(dmem.view(E * R, DF), dmp.view(E * R, DF), rnd=True).view(1, E * R, DF)
    sink.linear("img_feature_embedding", dmem_tot, cache["memory_r"].view(1, E * R, -1))
    dmemory = ops.matmul(dmem_tot, W.bwd("img_feature_embedding.weight")).view(E, R, -1)
    dtp = ops.copy2d_(ops.empty(E, Qp * DF), dt.view(E, Q * DF)[:, :Qp * DF], rnd=True).view(1, E * Qp, DF)
    sink.linear("prediction_embedding", dtp, cache["preds"].view(1, E * Qp, -1))
    dpreds = ops.matmul(dtp, W.bwd("prediction_embedding.weight"), rnd=True).view(E * Qp, -1)
    return dmemory, dpreds


# ------------------------------------------------------------------------------------- fusion A
def fusion_a_forward(ops, W, memory_r, preds, E, S, L, need_cache=True, want_aux_heads=False, drop=None):
    """-> dict(loss_vec [E,S*50], learned_loss [E], actions [E,4,4] (+ pred_boxes/pred_logits)), cache.
    Works for S in 1..5 (the policy rollout feeds 1-4 frames, reference models/interactron.py:174-197)."""
    R, Qp = S * L, S * NP
    Tn = R + Qp + NA
    HD = DF // NH
    seq = ops.empty(E, Tn, DF)
    lin(ops, memory_r.view(E, R, -1), W.w("img_feature_embedding.weight"), W.p("img_feature_embedding.bias"),
        out=seq[:, :R])
    lin(ops, preds.view(E, Qp, -1), W.w("prediction_embedding.weight"), W.p("prediction_embedding.bias"),
        out=seq[:, R:R + Qp])
    at = W.p("action_tokens")
    ops.copy2d_(seq.view(E, Tn * DF)[:, (R + Qp) * DF:], at.reshape(1, NA * DF).expand(E, NA * DF))
    spe = W.p("model.seq_pos_embed").reshape(1, -1, DF)[:, :Tn].reshape(1, Tn * DF)
    x = ops.add(seq.view(E, Tn * DF), spe.contiguous()).view(1, E * Tn, DF)             # residual stream fp32
    k_embd = _next(drop, "embd")                                       # train(): x = drop(seq + pos) (gpt.py:195)
    if drop is not None:
        x = ops.dropout(x, k_embd, out=x)
    caches = []
    for i in range(N_LAYERS):
        pre = f"model.blocks.{i}."
        h, h_r, m1, r1 = ops.layernorm_fwd(x.view(E * Tn, DF), W.p(pre + "ln1.weight"), W.p(pre + "ln1.bias"))
        h_r = h_r.view(1, E * Tn, DF)
        q = lin(ops, h_r, W.w(pre + "attn.query.weight"), W.p(pre + "attn.query.bias"), rnd=True)
        k = lin(ops, h_r, W.w(pre + "attn.key.weight"), W.p(pre + "attn.key.bias"), rnd=True)
        v = lin(ops, h_r, W.w(pre + "attn.value.weight"), W.p(pre + "attn.value.bias"), rnd=True)
        q, k, v = q.view(E, Tn, DF), k.view(E, Tn, DF), v.view(E, Tn, DF)
        # the reference's attention mask is all ones (models/gpt.py:35-36): full attention
        ka, kr = _next(drop, "attn"), _next(drop)                      # attn_drop (gpt.py:51), resid_drop (gpt.py:56)
        o, P = attention_fwd(ops, q, k, v, E, Tn, Tn, NH, HD, 1.0 / (HD ** 0.5), None, drop=ka)
        if drop is None:
            x1 = lin(ops, o.view(1, E * Tn, DF), W.w(pre + "attn.proj.weight"), W.p(pre + "attn.proj.bias"), residual=x)
        else:
            x1 = _drop_res(ops, lin(ops, o.view(1, E * Tn, DF), W.w(pre + "attn.proj.weight"),
                                    W.p(pre + "attn.proj.bias")), kr, x)
        h2, h2_r, m2, r2 = ops.layernorm_fwd(x1.view(E * Tn, DF), W.p(pre + "ln2.weight"), W.p(pre + "ln2.bias"))
        upre = ops.empty(1, E * Tn, 4 * DF)
        u = lin(ops, h2_r.view(1, E * Tn, DF), W.w(pre + "mlp.0.weight"), W.p(pre + "mlp.0.bias"), act="gelu",
                out_pre=upre, rnd=True)
        km = _next(drop)                                       # the Dropout closing the MLP (gpt.py:72)
        if drop is None:
            x2 = lin(ops, u, W.w(pre + "mlp.2.weight"), W.p(pre + "mlp.2.bias"), residual=x1)
        else:
            x2 = _drop_res(ops, lin(ops, u, W.w(pre + "mlp.2.weight"), W.p(pre + "mlp.2.bias")), km, x1)
        if need_cache:
            caches.append(dict(x=x, m1=m1, r1=r1, q=q, k=k, v=v, P=P, x1=x1, m2=m2, r2=r2, upre=upre,
                               h_r=h_r, o=o, h2_r=h2_r, u=u, drop=None if drop is None else (kr, km)))
        x = x2
    yf, yf_r, mf, rf = ops.layernorm_fwd(x.view(E * Tn, DF), W.p("model.ln_f.weight"), W.p("model.ln_f.bias"))
    yf_r = yf_r.view(E, Tn, DF)
    # head (no bias) only on the tokens that are decoded: predictions and the first 4 action tokens
    hw = W.w("model.head.weight")
    yp_in = ops.copy2d_(ops.empty(E, Qp * DF), yf_r.view(E, Tn * DF)[:, R * DF:(R + Qp) * DF]).view(E, Qp, DF)
    ya_in = ops.copy2d_(ops.empty(E, 4 * DF), yf_r.view(E, Tn * DF)[:, (R + Qp) * DF:(R + Qp + 4) * DF]).view(E, 4, DF)
    yp_r = lin(ops, yp_in, hw, rnd=True)
    ya_r = lin(ops, ya_in, hw, rnd=True)
    lv, norm, dlv, lhid = _learned_loss(ops, W, yp_r, E, S, need_cache)
    actions, ahid = mlp_fwd(ops, W, "action_decoder", ya_r)
    out = dict(loss_vec=lv, learned_loss=norm, actions=actions)
    if want_aux_heads:
        # direct-supervision heads, used by the detr_multiframe baseline only
        # (reference models/transformer.py:60-61, models/detr_multiframe.py:46-51)
        zb, _ = mlp_fwd(ops, W, "box_decoder", yp_r)
        out["pred_boxes"] = ops.sigmoid(zb)
        out["pred_logits"] = lin(ops, yp_r, W.w("logit_decoder.weight"), W.p("logit_decoder.bias"))
    cache = None
    if need_cache:
        cache = dict(layers=caches, x_last=x, mf=mf, rf=rf, yp_in=yp_in, yp_r=yp_r, dlv=dlv, lhid=lhid,
                     E=E, S=S, L=L, ya_in=ya_in, ya_r=ya_r, ahid=ahid, memory_r=memory_r, preds=preds, k_embd=k_embd)
    return out, cache


def fusion_a_backward(ops, W, cache, sink=None, dactions=None):
    """Gradient of sum_e learned_loss[e] (+ <dactions, actions> when given) wrt the inputs:
    -> (dmemory [E,S*L,256] fp32, dpreds [E*S*50,1496] TF32-clean).  `sink` (default: none, the
    inner loop) receives the gradients of the fusion parameters phi (meta-training step)."""
    E, S, L = cache["E"], cache["S"], cache["L"]
    R, Qp = S * L, S * NP
    Tn = R + Qp + NA
    HD = DF // NH
    sink = sink if sink is not None else NullSink()
    dz = ops.round_tf32(cache["dlv"].view(E, Qp, 1))
    dyp = mlp_bwd(ops, W, "loss_decoder", dz, cache["yp_r"], cache["lhid"], sink, rnd=True)  # d(head out) [E,Qp,512]
    sink.linear("model.head", dyp, cache["yp_in"])
    dyp_in = ops.matmul(dyp, W.bwd("model.head.weight"))                                  # d(ln_f out)
    dyf = ops.zeros(E, Tn, DF)
    ops.copy2d_(dyf.view(E, Tn * DF)[:, R * DF:(R + Qp) * DF], dyp_in.view(E, Qp * DF))
    if dactions is not None:
        dya = mlp_bwd(ops, W, "action_decoder", dactions, cache["ya_r"], cache["ahid"], sink, rnd=True)
        sink.linear("model.head", dya, cache["ya_in"], accumulate=True)
        dya_in = ops.matmul(dya, W.bwd("model.head.weight"))
        ops.copy2d_(dyf.view(E, Tn * DF)[:, (R + Qp) * DF:(R + Qp + 4) * DF], dya_in.view(E, 4 * DF))
    dx, _ = ops.layernorm_bwd(dyf.view(E * Tn, DF), cache["x_last"].view(E * Tn, DF), cache["mf"], cache["rf"],
                              W.p("model.ln_f.weight"), **sink.norm("model.ln_f"))
    for i in reversed(range(N_LAYERS)):
        pre = f"model.blocks.{i}."
        s = cache["layers"][i]
        dx_r = ops.round_tf32(dx).view(1, E * Tn, DF)
        dk_ = s.get("drop")
        if dk_ is not None:                   # gradient of the dropped MLP branch; the residual path keeps dx
            dx_r = ops.dropout(dx.view(1, E * Tn, DF), dk_[1])
        sink.linear(pre + "mlp.2", dx_r, s["u"])
        du = ops.matmul(dx_r, W.bwd(pre + "mlp.2.weight"), epi="gelu_grad", aux=s["upre"], rnd=True)
        sink.linear(pre + "mlp.0", du, s["h2_r"].view(1, E * Tn, DF))
        dh2 = ops.matmul(du, W.bwd(pre + "mlp.0.weight"))
        dx1, dx1_r = ops.layernorm_bwd(dh2.view(E * Tn, DF), s["x1"].view(E * Tn, DF), s["m2"], s["r2"],
                                       W.p(pre + "ln2.weight"), **sink.norm(pre + "ln2"))
        dx1 = ops.add(dx1, dx)                                                          # + residual path
        dx1_r = ops.round_tf32(dx1).view(1, E * Tn, DF)
        if dk_ is not None:
            dx1_r = ops.dropout(dx1.view(1, E * Tn, DF), dk_[0])
        sink.linear(pre + "attn.proj", dx1_r, s["o"].view(1, E * Tn, DF))
        dO = ops.matmul(dx1_r, W.bwd(pre + "attn.proj.weight"), rnd=True)
        dq, dk, dv = ops.empty(E, Tn, DF), ops.empty(E, Tn, DF), ops.empty(E, Tn, DF)
        attention_bwd(ops, dO.view(E, Tn, DF), s["q"], s["k"], s["v"], s["P"], E, Tn, Tn, NH, HD,
                      1.0 / (HD ** 0.5), dq, dk, dv)
        for nm, d_ in (("query", dq), ("key", dk), ("value", dv)):
            sink.linear(pre + "attn." + nm, d_.view(1, E * Tn, DF), s["h_r"])
        dh = ops.matmul(dq.view(1, E * Tn, DF), W.w(pre + "attn.query.weight"))
        ops.matmul(dk.view(1, E * Tn, DF), W.w(pre + "attn.key.weight"), out=dh, accumulate=True)
        ops.matmul(dv.view(1, E * Tn, DF), W.w(pre + "attn.value.weight"), out=dh, accumulate=True)
        dxa, _ = ops.layernorm_bwd(dh.view(E * Tn, DF), s["x"].view(E * Tn, DF), s["m1"], s["r1"],
                                   W.p(pre + "ln1.weight"), **sink.norm(pre + "ln1"))
        dx = ops.add(dxa, dx1)
    if cache.get("k_embd") is not None:       # through x = drop(seq + pos)
        dx = ops.dropout(dx.view(1, E * Tn, DF), cache["k_embd"]).view(E * Tn, DF)
    dseq = dx.view(E, Tn * DF)
    if sink.wants("model.seq_pos_embed"):
        assert Tn * DF == W.p("model.seq_pos_embed").numel(), "phi gradients need the full 5-frame sequence"
        sink.colsum("model.seq_pos_embed", dx.view(1, E, Tn * DF))
    sink.colsum("action_tokens", dx.view(1, E, Tn * DF)[:, :, (R + Qp) * DF:])
    dimg = ops.copy2d_(ops.empty(E, R * DF), dseq[:, :R * DF], rnd=True).view(1, E * R, DF)
    sink.linear("img_feature_embedding", dimg, cache["memory_r"].view(1, E * R, -1))
    dmemory = ops.matmul(dimg, W.bwd("img_feature_embedding.weight")).view(E, R, -1)
    dpe = ops.copy2d_(ops.empty(E, Qp * DF), dseq[:, R * DF:(R + Qp) * DF], rnd=True).view(1, E * Qp, DF)
    sink.linear("prediction_embedding", dpe, cache["preds"].view(1, E * Qp, -1))
    dpreds = ops.matmul(dpe, W.bwd("prediction_embedding.weight"), rnd=True).view(E * Qp, -1)
    return dmemory, dpreds
