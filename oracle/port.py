"""ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product path.

CPU restatement (plain PyTorch ops + autograd, fp32) of the reference's algorithm for the hot
path: DETR forward, fusion A / fusion B forward, learned loss, inner gradient by autograd,
clipped SGD step, re-detection.  It exists because the reference itself (a Python tree at
/root/reference) cannot travel to the GPU box: this file can, and is what `smoke()`, the GPU
tests' live cross-checks and bench.py's `cpu_baseline` / `--impl reference` legs run there.

Pinned: tests/test_oracle_port.py checks it here against the UNMODIFIED reference
(`oracle/reference_harness.py`) on the same weights/episodes, and against the committed goldens
in tests/golden/ that tools/make_golden.py produced from the reference.  Mode D1 (frozen
backbone) throughout, like every parity statement in this repo.

Only tests/, tools/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.
Functions operate on a flat `state_dict` (keys as in the reference checkpoints) — no nn.Module
forward of the product package is involved.
"""
import math

import torch
import torch.nn.functional as F

# ---------------------------------------------------------------------------------- pieces


def sine_position(mask, feats=128, temperature=10000.0):
    """reference models/detr_models/position_encoding.py:28-48 (normalize=True, scale=2*pi).
    mask bool [N,h,w] -> [N, 2*feats, h, w]."""
    nm = ~mask
    y = nm.cumsum(1, dtype=torch.float32)
    x = nm.cumsum(2, dtype=torch.float32)
    y = y / (y[:, -1:, :] + 1e-6) * (2 * math.pi)
    x = x / (x[:, :, -1:] + 1e-6) * (2 * math.pi)
    dim_t = torch.arange(feats, dtype=torch.float32, device=mask.device)
    dim_t = temperature ** (2 * (dim_t // 2) / feats)
    px, py = x[..., None] / dim_t, y[..., None] / dim_t
    px = torch.stack((px[..., 0::2].sin(), px[..., 1::2].cos()), dim=4).flatten(3)
    py = torch.stack((py[..., 0::2].sin(), py[..., 1::2].cos()), dim=4).flatten(3)
    return torch.cat((py, px), dim=3).permute(0, 3, 1, 2)


def mha(P, pre, q_in, k_in, v_in, nh, key_padding_mask=None):
    """nn.MultiheadAttention forward as the reference uses it (seq-first inputs [L,N,D];
    reference models/detr_models/transformer.py:154-155,219-226): separate q/k/v projections,
    q pre-scaled by 1/sqrt(hd), additive -inf key-padding mask, softmax, out_proj."""
    Lq, N, Dm = q_in.shape
    Lk = k_in.shape[0]
    hd = Dm // nh
    w, b = P[pre + "in_proj_weight"], P[pre + "in_proj_bias"]
    q = F.linear(q_in, w[:Dm], b[:Dm])
    k = F.linear(k_in, w[Dm:2 * Dm], b[Dm:2 * Dm])
    v = F.linear(v_in, w[2 * Dm:], b[2 * Dm:])
    q = q.reshape(Lq, N * nh, hd).transpose(0, 1) * (1.0 / math.sqrt(hd))
    k = k.reshape(Lk, N * nh, hd).transpose(0, 1)
    v = v.reshape(Lk, N * nh, hd).transpose(0, 1)
    s = torch.bmm(q, k.transpose(1, 2))
    if key_padding_mask is not None:
        add = torch.zeros(N, Lk, device=s.device).masked_fill(key_padding_mask, float("-inf"))
        s = s + add[:, None, None, :].expand(N, nh, 1, Lk).reshape(N * nh, 1, Lk)
    a = torch.softmax(s, dim=-1)
    o = torch.bmm(a, v).transpose(0, 1).reshape(Lq, N, Dm)
    return F.linear(o, P[pre + "out_proj.weight"], P[pre + "out_proj.bias"])


def ln(P, pre, x):
    return F.layer_norm(x, (x.shape[-1],), P[pre + "weight"], P[pre + "bias"], 1e-5)


def mlp(P, pre, x, n=3):
    """reference models/detr_models/detr.py:299-311."""
    for i in range(n):
        x = F.linear(x, P[f"{pre}layers.{i}.weight"], P[f"{pre}layers.{i}.bias"])
        if i < n - 1:
            x = F.relu(x)
    return x


def decoder_layer(P, pre, tgt, memory, pos, qpos, nh, kpm):
    """post-norm DETR decoder layer, reference models/detr_models/transformer.py:211-232."""
    q = tgt + qpos
    tgt = ln(P, pre + "norm1.", tgt + mha(P, pre + "self_attn.", q, q, tgt, nh))
    tgt = ln(P, pre + "norm2.", tgt + mha(P, pre + "multihead_attn.", tgt + qpos, memory + pos, memory, nh, kpm))
    ff = F.linear(F.relu(F.linear(tgt, P[pre + "linear1.weight"], P[pre + "linear1.bias"])),
                  P[pre + "linear2.weight"], P[pre + "linear2.bias"])
    return ln(P, pre + "norm3.", tgt + ff)


# ---------------------------------------------------------------------------------- DETR


def backbone_features(body, frames, masks):
    """Frozen trunk (decision D1) -> src [N,2048,h,w], bool mask [N,h,w]
    (reference models/detr_models/backbone.py:72-80)."""
    with torch.no_grad():
        src = body(frames)["0"]
    m = F.interpolate(masks[None].float(), size=src.shape[-2:]).to(torch.bool)[0]
    return src, m


def detr_forward(P, src, mask, pre="detector."):
    """reference models/detr_models/detr.py:66-75 + transformer.py:46-58 from backbone features."""
    N, _, h, w = src.shape
    pos = sine_position(mask).flatten(2).permute(2, 0, 1)
    x = F.conv2d(src, P[pre + "input_proj.weight"], P[pre + "input_proj.bias"]).flatten(2).permute(2, 0, 1)
    kpm = mask.flatten(1)
    for i in range(6):
        lp = f"{pre}transformer.encoder.layers.{i}."
        q = x + pos
        x = ln(P, lp + "norm1.", x + mha(P, lp + "self_attn.", q, q, x, 8, kpm))
        ff = F.linear(F.relu(F.linear(x, P[lp + "linear1.weight"], P[lp + "linear1.bias"])),
                      P[lp + "linear2.weight"], P[lp + "linear2.bias"])
        x = ln(P, lp + "norm2.", x + ff)
    memory = x
    qpos = P[pre + "query_embed.weight"][:, None, :].repeat(1, N, 1)
    tgt = torch.zeros_like(qpos)
    for j in range(6):
        tgt = decoder_layer(P, f"{pre}transformer.decoder.layers.{j}.", tgt, memory, pos, qpos, 8, kpm)
    hs = ln(P, pre + "transformer.decoder.norm.", tgt).transpose(0, 1)
    return {"pred_logits": F.linear(hs, P[pre + "class_embed.weight"], P[pre + "class_embed.bias"]),
            "pred_boxes": mlp(P, pre + "bbox_embed.", hs).sigmoid(),
            "image_features": src,
            "embedded_memory_features": memory.permute(1, 2, 0).reshape(N, -1, h, w),
            "box_features": hs}


# ---------------------------------------------------------------------------------- fusion


def _tokens(P, det, pre="fusion."):
    """Embeddings shared by both fusion nets (reference models/transformer.py:49-51)."""
    mem = det["embedded_memory_features"][None].permute(0, 1, 3, 4, 2)
    img = F.linear(mem, P[pre + "img_feature_embedding.weight"], P[pre + "img_feature_embedding.bias"])
    preds = torch.cat((det["box_features"], det["pred_logits"], det["pred_boxes"]), dim=-1)[None]
    pe = F.linear(preds, P[pre + "prediction_embedding.weight"], P[pre + "prediction_embedding.bias"])
    return img.reshape(1, -1, img.shape[-1]), pe.reshape(1, -1, pe.shape[-1])


def fusion_a(P, det, pre="fusion.", aux_heads=False):
    """reference models/transformer.py:47-66 + models/gpt.py:39-78,189-200 (mask all ones -> full
    attention; dropout off in eval)."""
    img, pe = _tokens(P, det, pre)
    n_pred = pe.shape[1]
    x = torch.cat((img, pe, P[pre + "action_tokens"]), dim=1)
    T = x.shape[1]
    x = x + P[pre + "model.seq_pos_embed"][:, :T]
    for i in range(4):
        bp = f"{pre}model.blocks.{i}."
        h = ln(P, bp + "ln1.", x)
        B, _, C = h.shape
        split = lambda t: t.view(B, T, 8, C // 8).transpose(1, 2)
        q = split(F.linear(h, P[bp + "attn.query.weight"], P[bp + "attn.query.bias"]))
        k = split(F.linear(h, P[bp + "attn.key.weight"], P[bp + "attn.key.bias"]))
        v = split(F.linear(h, P[bp + "attn.value.weight"], P[bp + "attn.value.bias"]))
        att = torch.softmax((q @ k.transpose(-2, -1)) * (1.0 / math.sqrt(C // 8)), dim=-1)
        y = (att @ v).transpose(1, 2).contiguous().view(B, T, C)
        x = x + F.linear(y, P[bp + "attn.proj.weight"], P[bp + "attn.proj.bias"])
        h2 = ln(P, bp + "ln2.", x)
        x = x + F.linear(F.gelu(F.linear(h2, P[bp + "mlp.0.weight"], P[bp + "mlp.0.bias"])),
                         P[bp + "mlp.2.weight"], P[bp + "mlp.2.bias"])
    y = F.linear(ln(P, pre + "model.ln_f.", x), P[pre + "model.head.weight"])
    yp = y[:, -(n_pred + 5):-5]
    out = {"loss": mlp(P, pre + "loss_decoder.", yp), "actions": mlp(P, pre + "action_decoder.", y[:, -5:-1])[0]}
    if aux_heads:
        out["pred_boxes"] = mlp(P, pre + "box_decoder.", yp).sigmoid()[0]
        out["pred_logits"] = F.linear(yp, P[pre + "logit_decoder.weight"], P[pre + "logit_decoder.bias"])[0]
    return out


def fusion_b(P, det, pre="fusion."):
    """reference models/new_transformer.py:34-58 (full 5-frame episodes, all-False key mask)."""
    img, pe = _tokens(P, det, pre)
    tgt = torch.cat((pe, P[pre + "action_tokens"]), dim=1).permute(1, 0, 2)
    memory = img.permute(1, 0, 2)
    pos = P[pre + "pos_embed"].permute(1, 0, 2)
    qpos = P[pre + "query_embed"].permute(1, 0, 2)
    for j in range(4):
        tgt = decoder_layer(P, f"{pre}transformer.layers.{j}.", tgt, memory, pos, qpos, 8, None)
    y = ln(P, pre + "transformer.norm.", tgt).permute(1, 0, 2)
    return {"loss": mlp(P, pre + "loss_decoder.", y[:, :-5]), "actions": mlp(P, pre + "action_decoder.", y[:, -5:-1])[0]}


# ---------------------------------------------------------------------------------- the path


def fast_weight_names(state_dict, pre="detector."):
    """theta in the reference's order (utils/meta_utils.py:5-24 under D1): parameters of leaf
    modules only, i.e. everything in the detector except in_proj_* and the frozen backbone."""
    return [k for k in state_dict
            if k.startswith(pre) and "backbone" not in k and "in_proj" not in k and "running_" not in k]


def predict(state_dict, body, data, kind, lr=1e-3, clip=0.01, trace=None):
    """reference models/interactron.py:31-59 / models/interactron_random.py:27-55 for one episode.
    kind "A" = fusion A (interactron), "B" = fusion B (interactron_random)."""
    frames, masks = data["frames"][0], data["masks"][0]
    src, m = backbone_features(body, frames, masks)
    names = fast_weight_names(state_dict)
    P = dict(state_dict)
    theta = [state_dict[n].detach().clone().requires_grad_(True) for n in names]
    P.update(zip(names, theta))
    det = detr_forward(P, src, m)
    fo = fusion_a(P, det) if kind == "A" else fusion_b(P, det)
    learned = torch.norm(fo["loss"])                                        # interactron.py:50
    grads = torch.autograd.grad(learned, theta, allow_unused=True)          # interactron.py:51-52
    fast = [p if g is None else p - torch.clip(lr * g, min=-clip, max=clip)  # meta_utils.py:135-142
            for p, g in zip(theta, grads)]
    P2 = dict(state_dict)
    P2.update(zip(names, [f.detach() for f in fast]))
    with torch.no_grad():
        post = detr_forward(P2, src[0:1], m[0:1])
    if trace is not None:
        trace.update(pre=det, learned_loss=learned.detach(), grads=grads, theta_prime=fast, names=names,
                     actions=fo["actions"].detach(), loss_vec=fo["loss"].detach().reshape(-1))
    return {k: v.detach().unsqueeze(0) for k, v in post.items()}


def detr_predict(state_dict, body, data, pre="detector."):
    """reference models/detr.py:20-40 (all b*s frames, no adaptation)."""
    b, s = data["frames"].shape[:2]
    src, m = backbone_features(body, data["frames"].flatten(0, 1), data["masks"].flatten(0, 1))
    with torch.no_grad():
        out = detr_forward(state_dict, src, m, pre)
    return {k: v.view(b, s, *v.shape[1:]) for k, v in out.items()}


def hungarian_cost(logits, boxes, tgt_labels, tgt_boxes, w_class=1.0, w_bbox=5.0, w_giou=2.0):
    """Cost matrix of reference models/detr_models/matcher.py:53-71 for one frame:
    logits [Q,C], boxes [Q,4] cxcywh, targets [T], [T,4] -> [Q,T]."""
    prob = logits.softmax(-1)

    def xyxy(b):
        cx, cy, w, h = b.unbind(-1)
        return torch.stack((cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h), -1)

    a, t = xyxy(boxes), xyxy(tgt_boxes)
    area_a = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    area_t = (t[:, 2] - t[:, 0]) * (t[:, 3] - t[:, 1])
    wh = (torch.min(a[:, None, 2:], t[:, 2:]) - torch.max(a[:, None, :2], t[:, :2])).clamp(min=0)
    inter = wh[..., 0] * wh[..., 1]
    union = area_a[:, None] + area_t - inter
    whc = (torch.max(a[:, None, 2:], t[:, 2:]) - torch.min(a[:, None, :2], t[:, :2])).clamp(min=0)
    areac = whc[..., 0] * whc[..., 1]
    giou = inter / union - (areac - union) / areac
    return w_bbox * torch.cdist(boxes, tgt_boxes, p=1) - w_class * prob[:, tgt_labels] - w_giou * giou


def _xyxy(b):
    cx, cy, w, h = b.unbind(-1)
    return torch.stack((cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h), -1)


def hungarian_match(logits, boxes, targets, w_class=1.0, w_bbox=5.0, w_giou=2.0):
    """reference models/detr_models/matcher.py:32-77: per-frame LSAP on the cost matrix ->
    [(index_i, index_j)] int64."""
    from scipy.optimize import linear_sum_assignment
    out = []
    for f, t in enumerate(targets):
        if t["labels"].numel() == 0:
            e = torch.zeros(0, dtype=torch.int64)
            out.append((e, e))
            continue
        c = hungarian_cost(logits[f].detach(), boxes[f].detach(), t["labels"], t["boxes"], w_class, w_bbox, w_giou)
        i, j = linear_sum_assignment(c.numpy())
        out.append((torch.as_tensor(i, dtype=torch.int64), torch.as_tensor(j, dtype=torch.int64)))
    return out


def set_criterion(logits, boxes, targets, indices, background_c=0.1):
    """reference models/detr_models/detr.py:220-265 with losses = labels, boxes, cardinality
    (:111-167), given the matcher's indices.  logits [F,Q,C], boxes [F,Q,4] -> dict of 0-dim tensors
    (differentiable wrt logits / boxes)."""
    F_, Q, C = logits.shape
    bi = torch.cat([torch.full_like(i, f) for f, (i, _) in enumerate(indices)])
    si = torch.cat([i for i, _ in indices])
    tco = torch.cat([t["labels"][j] for t, (_, j) in zip(targets, indices)])
    tc = torch.full((F_, Q), C - 1, dtype=torch.int64)
    tc[bi, si] = tco                                                            # :118-122
    w = torch.ones(C, dtype=logits.dtype)
    w[-1] = background_c                                                        # :124-125
    logp = torch.log_softmax(logits, -1)
    nll = -logp.gather(-1, tc[..., None])[..., 0]
    loss_ce = (w[tc] * nll).sum() / w[tc].sum()                                 # F.cross_entropy(weight) :126
    if tco.numel():
        class_error = 100 - 100.0 * (logits[bi, si].argmax(-1) == tco).float().mean()   # :131, misc.py:431-446
    else:
        class_error = torch.tensor(100.0)
    lens = torch.tensor([float(t["labels"].numel()) for t in targets])
    card = (logits.argmax(-1) != C - 1).sum(1).float()
    cardinality_error = (card - lens).abs().mean()                              # :143-145
    num_boxes = max(float(lens.sum()), 1.0)                                     # :238-242 (single process)
    src = boxes[bi, si]
    tgt = torch.cat([t["boxes"][j] for t, (_, j) in zip(targets, indices)], 0)
    loss_bbox = (src - tgt).abs().sum() / num_boxes                             # :157-160
    a, t = _xyxy(src), _xyxy(tgt)
    area_a = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    area_t = (t[:, 2] - t[:, 0]) * (t[:, 3] - t[:, 1])
    wh = (torch.min(a[:, 2:], t[:, 2:]) - torch.max(a[:, :2], t[:, :2])).clamp(min=0)
    inter = wh[:, 0] * wh[:, 1]
    union = area_a + area_t - inter
    whc = (torch.max(a[:, 2:], t[:, 2:]) - torch.min(a[:, :2], t[:, :2])).clamp(min=0)
    areac = whc[:, 0] * whc[:, 1]
    giou = inter / union - (areac - union) / areac                              # util/box_ops.py:38-58 (diagonal)
    loss_giou = (1 - giou).sum() / num_boxes                                    # :162-165
    return {"loss_ce": loss_ce, "class_error": class_error, "loss_bbox": loss_bbox, "loss_giou": loss_giou,
            "cardinality_error": cardinality_error}


# ------------------------------------------------------------------ evaluator post-processing
def match_predictions_to_detections(ious):
    """Restatement of utils/detection_utils.py:401-421 (proposal rounds between predictions and ground
    truths) on a CPU tensor [p, g] -> (best_ious [g], best_idxs [g])."""
    p, g = ious.shape
    dev = ious.device                                                           # the reference allocates on ious.device
    pref = torch.argsort(ious, dim=1, descending=True)                           # :402
    ptr = torch.zeros(p, dtype=torch.long, device=dev)
    free = torch.ones(p, device=dev).bool()
    tent = -torch.ones(g, dtype=torch.long, device=dev)
    for _ in range(g):                                                          # :406
        prop = pref[torch.arange(p, device=dev), ptr]
        for j in range(g):
            new = torch.argmax(ious[:, j] * (prop == j))                         # :409
            if tent[j] != -1 and tent[j] != new:
                free[tent[j]] = True
            tent[j] = new
            free[tent[j]] = False
        ptr[free] += 1                                                          # :414
        if torch.count_nonzero(~free) >= min(p, g):
            break
    best = torch.zeros(g, device=dev)
    best[tent != -1] = ious[tent[tent != -1], tent != -1]                         # :419
    tent[best == 0.0] = -1
    return best, tent


def evaluator_records(pred_logits, pred_boxes, gt_boxes_cxcywh, gt_cats, img, class_ids, match_fn=None,
                      background=1235):
    """Restatement of the per-image body of engine/random_policy_evaluator.py:63-146 with torch /
    torchvision CPU ops: pred_logits [Q,C], pred_boxes [Q,4] cxcywh -> list of TP / FP / FN dicts.
    match_fn: the reference's own match_predictions_to_detections when it is importable (golden
    generation), else the restatement above."""
    import torchvision
    match_fn = match_fn or match_predictions_to_detections
    item = lambda t: t.item()
    pb = _xyxy(pred_boxes)                                                       # :65
    ps, pc = pred_logits.softmax(dim=-1).max(dim=-1)                             # :66
    gb, gc = _xyxy(gt_boxes_cxcywh.reshape(-1, 4)), gt_cats
    keep = pc != background                                                      # :70-73
    pb, pc, ps = pb[keep], pc[keep], ps[keep]
    kept = torchvision.ops.nms(pb, ps, iou_threshold=0.5)                        # :75
    pc, pb, ps = pc[kept], pb[kept], ps[kept]
    pset, gset = set([int(c) for c in pc]), set([int(c) for c in gc])            # :80-81
    ponly = set(class_ids).intersection(pset - gset)

    def rec(kind, match, cat, iou, score, box):
        return {"iou": iou, "category_match": match, "type": kind, "pred_cat": cat, "pred_score": score,
                "box": [item(c) for c in box], "area": item((box[2] - box[0]) * (box[3] - box[1])), "img": img}

    out = []
    for cat in gset:                                                            # :84
        if torch.any(pc == cat):
            cpb, cps, cgb = pb[pc == cat], ps[pc == cat], gb[gc == cat]
            ious = torchvision.ops.box_iou(cpb, cgb)
            best_ious, best_idx = match_fn(ious)
            for i in range(ious.shape[0]):                                      # :91-114
                kind = "tp" if torch.any(best_idx == i) else "fp"
                out.append(rec(kind, True, cat, item(ious[i].max()), item(cps[i]), cpb[i]))
            for j in range(ious.shape[1]):                                      # :115-127
                if best_ious[j] == 0.0:
                    out.append(rec("fn", False, cat, 0.0, 0.0, cgb[j]))
        else:                                                                   # :128-141
            for box in gb[gc == cat]:
                out.append(rec("fn", False, cat, 0.0, 0.0, box))
    for cat in ponly:                                                           # :142-156
        for box, sc in zip(pb[pc == cat], ps[pc == cat]):
            out.append(rec("fp", False, cat, 0.0, item(sc), box))
    return out
