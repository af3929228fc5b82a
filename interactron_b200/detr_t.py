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

from .layers import (DecDims, GradSink, MultiSink, NullSink, T, attention_bwd, attention_fwd,  # noqa: F401
                     cross_kv_all, decoder_layer_bwd, decoder_layer_fwd, lin, ln_plus, mlp_bwd, mlp_fwd, _drop_res, _next)

D, H, HD, FFN, NQ = 256, 8, 32, 2048, 50
N_ENC, N_DEC = 6, 6
SCALE = 1.0 / math.sqrt(HD)


# --------------------------------------------------------------------------- forward
def detr_t_forward(ops, W, src_r, pos, kmask, E, Fe, L, preds=None, need_cache=True, drop=None):
    """
    src_r : [E, Fe*L, 2048] TF32-clean backbone features, token-major
    pos   : [E*Fe*L, 256] sine position embedding (full fp32)
    kmask : uint8 [E*Fe, L] (1 = padded key) or None
    preds : optional [E*Fe*50, 1496] buffer that receives cat(box_features, logits, boxes)
            TF32-clean, i.e. the prediction tokens fusion embeds (reference models/transformer.py:50)
    drop  : layers.DropCtx in train() mode (dropout p=0.1 after every attention softmax, sub-layer output and FFN
            activation, reference transformer.py:154-159,219-230), None in eval()
    Returns (out, cache): out has logits [E,Fe*50,C], boxes [E,Fe*50,4], hs, memory (+ `_r` twins).
    """
    R, Q, B = Fe * L, Fe * NQ, E * Fe
    c = {} if need_cache else None
    ip_w = W.w("input_proj.weight")
    ip_w = ip_w.reshape(ip_w.shape[0], D, -1)
    x = ops.empty(E, R, D)
    x_r = lin(ops, src_r, ip_w, W.p("input_proj.bias"), out_pre=x, rnd=True)        # x (fp32), x_r
    enc = []
    x_pos = None                     # x + pos of the current layer input (fused into the previous layer's norm2)
    for i in range(N_ENC):
        pre = f"transformer.encoder.layers.{i}."
        ipw, ipb = W.w(pre + "self_attn.in_proj_weight"), W.p(pre + "self_attn.in_proj_bias")
        qk_in = (x_pos if x_pos is not None else ops.add(x.view(E * R, D), pos, rnd=True)).view(1, E * R, D)
        qk = lin(ops, qk_in, ipw[:, :2 * D], ipb[:, :2 * D], rnd=True)               # [1,E*R,512]
        v = lin(ops, x_r.view(1, E * R, D), ipw[:, 2 * D:], ipb[:, 2 * D:], rnd=True)
        qk3, v3 = qk.view(B, L, 2 * D), v.view(B, L, D)
        # train(): keys in module order (transformer.py:154-159): attention probabilities, dropout1, FFN dropout, dropout2
        ka, kd1 = _next(drop, "attn"), _next(drop)
        o, P = attention_fwd(ops, qk3[..., :D], qk3[..., D:], v3, B, L, L, H, HD, SCALE, kmask, drop=ka)
        if drop is None:
            a = lin(ops, o.view(E, R, D), W.w(pre + "self_attn.out_proj.weight"),
                    W.p(pre + "self_attn.out_proj.bias"), residual=x)
        else:
            a = _drop_res(ops, lin(ops, o.view(E, R, D), W.w(pre + "self_attn.out_proj.weight"),
                                   W.p(pre + "self_attn.out_proj.bias")), kd1, x)
        x1, x1_r, m1, r1 = ops.layernorm_fwd(a.view(E * R, D), W.p(pre + "norm1.weight"),
                                             W.p(pre + "norm1.bias"))
        h = lin(ops, x1_r.view(E, R, D), W.w(pre + "linear1.weight"), W.p(pre + "linear1.bias"),
                act="relu", rnd=True)
        kf, kd2 = _next(drop), _next(drop)
        if drop is None:
            f = lin(ops, h, W.w(pre + "linear2.weight"), W.p(pre + "linear2.bias"), residual=x1.view(E, R, D))
        else:
            h = ops.dropout(h, kf, out=h)
            f = _drop_res(ops, lin(ops, h, W.w(pre + "linear2.weight"), W.p(pre + "linear2.bias")), kd2, x1.view(E, R, D))
        x2, x2_r, m2, r2, x_pos = ln_plus(ops, f.view(E * R, D), W.p(pre + "norm2.weight"), W.p(pre + "norm2.bias"),
                                          pos, (E * R, D))
        if need_cache:
            enc.append(dict(x_r=x_r, qk_in=qk_in, qk3=qk3, v3=v3, P=P, o=o, a=a, m1=m1, r1=r1,
                            x1_r=x1_r, h=h, f=f, m2=m2, r2=r2, drop=None if drop is None else (kd1, kf, kd2)))
        x, x_r = x2.view(E, R, D), x2_r.view(E, R, D)
    memory, memory_r = x, x_r
    mem_pos_r = x_pos.view(1, E * R, D)                          # memory + pos: the last norm2's second output

    qpos = W.p("query_embed.weight")                                                  # [Gw,50,256]
    dm = DecDims(E, B, NQ, L, D, H)
    tgt = ops.zeros(E, Q, D)
    tgt_r = tgt
    dec = []
    tq = None                        # tgt + query_pos, fused into the previous layer's norm3
    kvs = cross_kv_all(ops, W, [f"transformer.decoder.layers.{j}." for j in range(N_DEC)], dm, mem_pos_r,
                       memory_r.view(1, E * R, D))
    for j in range(N_DEC):
        res = decoder_layer_fwd(ops, W, f"transformer.decoder.layers.{j}.", dm, tgt, tgt_r, qpos,
                                mem_pos_r, memory_r.view(1, E * R, D), kmask, need_cache, drop=drop,
                                tgt_plus=tq, want_plus=j < N_DEC - 1, kv=None if kvs is None else kvs[j])
        tgt, tgt_r, dc = res[:3]
        tq = res[3] if j < N_DEC - 1 else None
        dec.append(dc)

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
    z, bb_hid = mlp_fwd(ops, W, "bbox_embed", hs_r)
    boxes = ops.sigmoid(z)
    if preds is not None:
        ops.copy2d_(preds[:, D + C:], boxes.view(E * Q, 4), rnd=True)
    out = dict(logits=logits, boxes=boxes, hs=hs, hs_r=hs_r, memory=memory, memory_r=memory_r)
    if need_cache:
        c.update(enc=enc, dec=dec, src_r=src_r, tgt_last=tgt, mh=mh, rh=rh, hs_r=hs_r, bb_hid=bb_hid,
                 boxes=boxes, E=E, Fe=Fe, L=L)
    return out, c


# --------------------------------------------------------------------------- backward
def detr_t_backward(ops, W, cache, sink, dpreds=None, dmemory=None, dlogits=None, dboxes=None, dhs=None):
    """Back-propagates into the flat gradient buffer(s) behind `sink`: the 157 fast weights on the
    inner loop (per-episode GradSink over theta); in the meta-training step also / only the shared
    `in_proj_*` parameters psi (reference models/interactron.py:123,134 - they stay live Parameters).

    Upstream gradients: either `dpreds` [E*Fe*50, 1496] (TF32-clean gradient of the fusion
    prediction tokens = cat(d box_features, d logits, d boxes)) and `dmemory` [E, Fe*L, 256]
    (gradient of the encoder memory from fusion), or explicit dlogits/dboxes/dhs.
    The backbone is frozen (D1).
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
    dhs_t = mlp_bwd(ops, W, "bbox_embed", dz, hs_r, cache["bb_hid"], sink, residual=dhs_in)
    sink.linear("class_embed", dlogits_r, hs_r)
    ops.matmul(dlogits_r, W.bwd("class_embed.weight"), out=dhs_t, accumulate=True)
    dt, _ = ops.layernorm_bwd(dhs_t.view(E * Q, D), cache["tgt_last"].view(E * Q, D), cache["mh"], cache["rh"],
                              W.p("transformer.decoder.norm.weight"), **sink.norm("transformer.decoder.norm"))

    # decoder -------------------------------------------------------------------------
    dmem = dmemory.contiguous() if dmemory is not None else ops.zeros(E, R, D)         # accumulates
    dmp = ops.zeros(1, E * R, D)                                                        # d(memory+pos)
    dqpos = ops.zeros(E, Q, D)                                                          # d(query_pos), per frame
    dm = DecDims(E, B, NQ, L, D, H)
    for j in reversed(range(N_DEC)):
        dt = decoder_layer_bwd(ops, W, f"transformer.decoder.layers.{j}.", dm, cache["dec"][j], dt, sink,
                               dqpos, dmp, dmem.view(1, E * R, D))
    # query_embed gradient: sum the per-frame query_pos gradients of each episode
    sink.colsum("query_embed.weight", dqpos.view(E, Fe, NQ * D))
    # gradient of memory = decoder V paths (dmem) + K paths through memory+pos (dmp) [+ fusion]
    dx = ops.add(dmem.view(E * R, D), dmp.view(E * R, D))

    # encoder -------------------------------------------------------------------------
    for i in reversed(range(N_ENC)):
        pre = f"transformer.encoder.layers.{i}."
        s = cache["enc"][i]
        ip_name, ip_bias = pre + "self_attn.in_proj_weight", pre + "self_attn.in_proj_bias"
        dk_ = s.get("drop")
        # without dropout the residual branch's gradient IS the LayerNorm's dx: its linear's bias gradient is fused
        df, df_r = ops.layernorm_bwd(dx, s["f"].view(E * R, D), s["m2"], s["r2"], W.p(pre + "norm2.weight"),
                                     **sink.norm(pre + "norm2", bias_of=pre + "linear2" if dk_ is None else None))
        df3, df3_r = df.view(E, R, D), df_r.view(E, R, D)
        kd1, kf, kd2 = dk_ if dk_ is not None else (None,) * 3
        if dk_ is not None:                   # gradient of the dropped branch; the residual path keeps df3
            dfm = ops.dropout(df3, kd2)
            dfm_r = dfm
        else:
            dfm, dfm_r = df3, df3_r
        sink.linear(pre + "linear2", dfm_r, s["h"], dfm)
        dh = ops.matmul(dfm_r, W.bwd(pre + "linear2.weight"), epi="relu_mask", aux=s["h"], rnd=True)
        if dk_ is not None:
            dh = ops.dropout(dh, kf, out=dh)  # h is stored dropped: relu_mask kept (h > 0) & keep, this adds 1/(1-p)
        sink.linear(pre + "linear1", dh, s["x1_r"].view(E, R, D))
        dx1 = ops.matmul(dh, W.bwd(pre + "linear1.weight"), residual=df3)
        da, da_r = ops.layernorm_bwd(dx1.view(E * R, D), s["a"].view(E * R, D), s["m1"], s["r1"],
                                     W.p(pre + "norm1.weight"),
                                     **sink.norm(pre + "norm1", bias_of=pre + "self_attn.out_proj" if dk_ is None else None))
        da3, da3_r = da.view(E, R, D), da_r.view(E, R, D)
        dam, dam_r = da3, da3_r
        if dk_ is not None:
            dam = ops.dropout(da3, kd1)
            dam_r = dam
        sink.linear(pre + "self_attn.out_proj", dam_r, s["o"].view(E, R, D), dam)
        dO = ops.matmul(dam_r, W.bwd(pre + "self_attn.out_proj.weight"), rnd=True)
        dqk, dv = ops.empty(B, L, 2 * D), ops.empty(B, L, D)
        attention_bwd(ops, dO.view(B, L, D), s["qk3"][..., :D], s["qk3"][..., D:], s["v3"], s["P"],
                      B, L, L, H, HD, SCALE, dqk[..., :D], dqk[..., D:], dv)
        if sink.wants(ip_name):
            sink.rows(ip_name, ip_bias, 0, 2 * D, dqk.view(1, E * R, 2 * D), s["qk_in"])
            sink.rows(ip_name, ip_bias, 2 * D, 3 * D, dv.view(1, E * R, D), s["x_r"].view(1, E * R, D))
        dxi = ops.matmul(dqk.view(1, E * R, 2 * D), W.bwd(ip_name, 0, 2 * D), residual=da.view(1, E * R, D))
        last = i == 0
        ops.matmul(dv.view(1, E * R, D), W.bwd(ip_name, 2 * D, 3 * D), out=dxi, accumulate=True, rnd=last)
        dx = dxi.view(E * R, D)

    # input_proj ----------------------------------------------------------------------
    dx3 = dx.view(E, R, D)                                            # TF32-clean (rounded by the last GEMM)
    sink.linear("input_proj", dx3, cache["src_r"])
