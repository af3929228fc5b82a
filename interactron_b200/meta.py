"""The meta-training step: `interactron.forward(data)` / `interactron_random.forward(data)`
(reference models/interactron.py:61-151, models/interactron_random.py:57-136; BASELINE config 5).

Per episode the reference does (theta = fast weights, psi = in_proj_*, phi = fusion):
  supervisor  g = d l(theta, psi, phi)/d theta with create_graph, theta' = theta - clip(lr*g),
              L_sup = criterion(detector_theta'(5 frames)) [+ CE(actions, best path)], backward:
              grads on phi and psi only (theta was detached), through theta' -> g (second order).
  detector    theta'' = clone(theta) - clip(lr * g.detach()), L_det = criterion(detector_theta''(1
              random frame)), backward: first-order grads on theta (identity Jacobian) and psi.
Gradients are *summed* over the episodes of the batch into `.grad`; the caller never calls backward.

Here, for all E episodes of the batch at once (every arithmetic step in the sm_100a kernels):
  1. pre-adapt forward + backward -> g [E,n]                         (the predict() path, cached)
  2. theta' and the clip mask in one kernel; W^T twins of theta'
  3. post-adapt forward on the 5 frames with per-episode theta'; matcher + criterion (+ gradient)
  4. its backward -> dL_sup/dtheta' [E,n] and the direct psi gradient
  5. v = -lr * mask * dL_sup/dtheta'; ONE dual-number pass of step 1 with theta-tangent v
     (interactron_b200/dual.py): the tangents of the phi / psi gradients are the second-order
     meta-gradients, and the policy loss's first-order gradient rides along as a tangent seed
  6. 1-frame post-adapt forward/backward for the detector loss (theta'' == theta' numerically)
In train() mode (both reference trainers call model.train()) every forward pass applies dropout with
counter-based masks (layers.DropCtx, csrc/itn_philox.cuh): the pre-adapt pass and the fusion network draw
their masks once and the backward of step 1 AND the dual pass of step 5 regenerate exactly those; the
post-adapt passes of steps 3 and 6 draw their own.  PyTorch's generator cannot be reproduced, so train()-mode
parity with the reference is statistical (tools/dropout_stats.py); eval() mode is the exact-parity mode.
"""
import random

import torch
import torch.nn.functional as F

from . import detr_t, fusion, parallel
from .dual import Dual, DualOps, DualWeights
from .layers import GradSink, MultiSink

LOSS_W = (1.0, 2.0, 5.0)            # ce + 2*bbox + 5*giou  (reference models/interactron.py:121-122,133)


class PathStorage:
    """Trie of action paths keeping, at every node, the first action of the cheapest path seen
    through it (reference utils/storage_utils.py:4-50: Node/Edge/PathStorage)."""

    def __init__(self):
        self.root = {"cost": float("inf"), "action": None, "next": {}}

    def add_path(self, path, cost):
        node = self.root
        for a in path:
            a = int(a)
            if cost < node["cost"]:
                node["cost"], node["action"] = cost, a
            node = node["next"].setdefault(a, {"cost": float("inf"), "action": None, "next": {}})

    def get_label(self, path):
        out, node = [], self.root
        for a in path:
            out.append(node["action"])
            node = node["next"][int(a)]
        return out


def _targets(data, E, S):
    return [{"labels": data["category_ids"][e][f], "boxes": data["boxes"][e][f]} for e in range(E) for f in range(S)]


class MetaParts:
    """The two launch-only halves of the step, separated by the host work (LSAP of the matcher, the
    path trie): `part_a` = steps 1-3 (+ the 1-frame pass of step 6) up to the detector outputs the
    criterion needs, `part_b` = steps 4-6 from the criterion's gradients to the flat meta-gradient.
    Both are pure kernel sequences on static shapes, so each is captured into a CUDA graph
    (`graph.GraphedCall`) and replayed; `part_b` reads the activations `part_a` left in `self.st`."""

    def __init__(self, model, train=False):
        self.model, self.loop = model, model._get_loop()
        self.train = bool(train)
        self.st = None

    def part_a(self, frames, masks, idx):
        loop = self.loop
        ops = loop.ops
        E, S = frames.shape[:2]
        C = loop.detector.class_embed.out_features
        NQ, D = detr_t.NQ, detr_t.D
        tpk = loop.theta_pack
        f_fwd = fusion.fusion_a_forward if loop.kind == "A" else fusion.fusion_b_forward
        f_bwd = fusion.fusion_a_backward if loop.kind == "A" else fusion.fusion_b_backward
        # 1. pre-adapt pass, learned loss, inner gradient
        src_r, pos, kmask, (h, w) = loop.features(frames.flatten(0, 1), masks.flatten(0, 1))
        L = h * w
        src3 = src_r.view(E, S * L, -1)
        Wd = loop._det_weights(loop.theta, loop.theta_r, loop.theta_t)
        Wf = loop._fusion_weights()
        preds = ops.empty(E * S * NQ, D + C + 4)
        tr = self.train
        pre, cache = detr_t.detr_t_forward(ops, Wd, src3, pos, kmask, E, S, L, preds=preds,
                                           drop=loop.drop_ctx("pre", tr))
        fout, fcache = f_fwd(ops, Wf, pre["memory_r"], preds, E, S, L, drop=loop.drop_ctx("fusion", tr))
        dmemory, dpreds = f_bwd(ops, Wf, fcache)
        g = ops.empty(E, tpk.numel)
        detr_t.detr_t_backward(ops, Wd, cache, GradSink(ops, tpk, g), dpreds=dpreds, dmemory=dmemory)
        del cache, fcache, dmemory, dpreds
        # 2. fast weights + clip mask, W^T twins of theta'
        theta_p, theta_p_r, cmask = ops.sgd_clip_update(loop.theta, g, loop.lr, loop.clip, want_mask=True)
        theta_p_t = tpk.transpose_into(ops, theta_p, ops.zeros(E, tpk.numel))
        Wp = loop._det_weights(theta_p, theta_p_r, theta_p_t)
        # 3. post-adapt pass on the 5 frames; 6a. on one chosen frame per episode (theta'' == theta')
        post, pcache = detr_t.detr_t_forward(ops, Wp, src3, pos, kmask, E, S, L, drop=loop.drop_ctx("post", tr))
        src1 = src_r.index_select(0, idx).view(E, L, -1)
        pos1 = pos.view(E * S, L, -1).index_select(0, idx).reshape(E * L, -1)
        km1 = kmask.index_select(0, idx)
        post1, c1 = detr_t.detr_t_forward(ops, Wp, src1, pos1, km1, E, 1, L, drop=loop.drop_ctx("post1", tr))
        self.st = dict(E=E, S=S, L=L, C=C, src3=src3, pos=pos, kmask=kmask, cmask=cmask, Wp=Wp, pcache=pcache, c1=c1)
        return {"post_logits": post["logits"].view(E * S, NQ, C), "post_boxes": post["boxes"].view(E * S, NQ, 4),
                "post1_logits": post1["logits"].view(E, NQ, C), "post1_boxes": post1["boxes"].view(E, NQ, 4),
                "actions": fout["actions"]}

    def part_b1(self, dlog, dbox, dact=None):
        """Steps 4-5.  Leaves the flat meta-gradient buffer G = [theta | psi | phi] in self.st with its phi
        segment FINAL (the second-order gradients of the fusion network) -> {"Gphi": that segment}."""
        loop, st = self.loop, self.st
        ops = loop.ops
        E, S, L, C = st["E"], st["S"], st["L"], st["C"]
        NQ, D = detr_t.NQ, detr_t.D
        tpk, ppk, fpk = loop.theta_pack, loop.psi_pack, loop.phi_pack
        n_t, n_p, n_f = tpk.numel, ppk.numel, fpk.numel
        f_fwd = fusion.fusion_a_forward if loop.kind == "A" else fusion.fusion_b_forward
        f_bwd = fusion.fusion_a_backward if loop.kind == "A" else fusion.fusion_b_backward
        Wp, src3, pos, kmask = st["Wp"], st["src3"], st["pos"], st["kmask"]
        # 4. backward of the post-adapt pass: dL_sup/dtheta' per episode, direct psi gradient
        g_sup = ops.empty(E, n_t)
        gpsi = ops.zeros(1, n_p)
        detr_t.detr_t_backward(ops, Wp, st["pcache"],
                               MultiSink(GradSink(ops, tpk, g_sup), GradSink(ops, ppk, gpsi, shared=True)),
                               dlogits=dlog.view(E, S * NQ, C), dboxes=dbox.view(E, S * NQ, 4))
        # 5. second order: one dual pass of step 1 along v = -lr * mask * dL_sup/dtheta'
        v = ops.mul_mask_u8(g_sup, st["cmask"], -loop.lr)
        v_t = tpk.transpose_into(ops, v, ops.zeros(E, n_t))
        dops = DualOps(ops)
        DWd = DualWeights((tpk, loop.theta, loop.theta_t, v, v_t), (ppk, loop.psi, loop.psi_t, None, None))
        DWf = DualWeights((fpk, loop.phi, loop.phi_t, None, None))
        preds2 = dops.empty(E * S * NQ, D + C + 4)
        # the dual pass re-runs step 1 on dual numbers: same sites, same seed -> the same dropout masks
        pre2, cache2 = detr_t.detr_t_forward(dops, DWd, src3, pos, kmask, E, S, L, preds=preds2,
                                             drop=loop.drop_ctx("pre", self.train))
        _, fcache2 = f_fwd(dops, DWf, pre2["memory_r"], preds2, E, S, L, drop=loop.drop_ctx("fusion", self.train))
        # the step's meta-gradient: ONE flat buffer [theta | psi | phi] (what gets all-reduced)
        G = ops.zeros(1, n_t + n_p + n_f)
        gphi2 = Dual(ops.zeros(1, n_f), G[:, n_t + n_p:])
        gpsi2 = dops.zeros(1, n_p)
        # the policy loss's first-order gradient rides along as a tangent seed on the action logits
        seed = None if dact is None else Dual(ops.zeros(E, 4, 4), dact)
        dmem2, dpreds2 = f_bwd(dops, DWf, fcache2, sink=GradSink(dops, fpk, gphi2, shared=True), dactions=seed)
        detr_t.detr_t_backward(dops, DWd, cache2, GradSink(dops, ppk, gpsi2, shared=True), dpreds=dpreds2,
                               dmemory=dmem2)
        del cache2, fcache2
        st["G"], st["gpsi"], st["gpsi2"] = G, gpsi, gpsi2.t
        return {"Gphi": G[:, n_t + n_p:]}

    def part_b2(self, dlog1, dbox1):
        """Step 6b: detector loss backward (first order): theta gets the gradient wrt theta'' unchanged.
        Completes the theta and psi segments of G -> {"Gtp": [theta | psi] segment}."""
        loop, st = self.loop, self.st
        ops = loop.ops
        E, C = st["E"], st["C"]
        NQ = detr_t.NQ
        tpk, ppk = loop.theta_pack, loop.psi_pack
        n_t, n_p = tpk.numel, ppk.numel
        G = st["G"]
        g_det = ops.empty(E, n_t)
        gpsi1 = ops.zeros(1, n_p)
        detr_t.detr_t_backward(ops, st["Wp"], st["c1"],
                               MultiSink(GradSink(ops, tpk, g_det), GradSink(ops, ppk, gpsi1, shared=True)),
                               dlogits=dlog1.view(E, NQ, C), dboxes=dbox1.view(E, NQ, 4))
        ops.colsum(g_det.view(1, E, n_t), out=G[:, :n_t])
        ops.copy2d_(G[:, n_t:n_t + n_p], ops.add(ops.add(st["gpsi"], st["gpsi2"]), gpsi1))
        return {"Gtp": G[:, :n_t + n_p]}

    def part_b(self, dlog, dbox, dlog1, dbox1, dact=None):
        self.part_b1(dlog, dbox, dact)
        self.part_b2(dlog1, dbox1)
        return {"G": self.st["G"]}


def _parts_runner(model, frames, masks, idx):
    """-> (run_a, run_b1, run_b2): the launch-only pieces, CUDA-graph replayed when the model allows it."""
    loop = model._get_loop()
    use_graph = model.use_cuda_graph and frames.is_cuda
    train = model.mode == "train"
    if not use_graph:
        parts = MetaParts(model, train)
        return parts.part_a, parts.part_b1, parts.part_b2
    from .graph import GraphedCall
    bb = loop.detector.backbone
    bb_key = tuple((t.data_ptr(), t._version) for t in list(bb.parameters()) + list(bb.buffers()))
    # everything the capture bakes in as a kernel scalar or a branch is part of the key
    key = ("meta", tuple(frames.shape), tuple(masks.shape), loop.kind, bb_key, float(loop.lr), float(loop.clip),
           loop.ops.precision_key, loop.backbone_impl, bool(loop.ops.fused_attention), train)
    ent = model._graphs.get(key)
    if ent is None:
        for k in [k for k in model._graphs if k[0] == "meta"]:
            del model._graphs[k]                       # one meta geometry at a time: the caches are large
        ent = model._graphs[key] = {"parts": MetaParts(model, train), "a": None, "b1": None, "b2": None}
    parts = ent["parts"]

    def run_a(f, m, i):
        if ent["a"] is None:
            ent["a"] = GraphedCall(parts.part_a, [f, m, i])
        return ent["a"](f, m, i, clone=False)

    def run_b1(*t):
        if ent["b1"] is None:
            ent["b1"] = GraphedCall(lambda *x: parts.part_b1(*x), list(t))
        return ent["b1"](*t, clone=False)

    def run_b2(*t):
        if ent["b2"] is None:
            ent["b2"] = GraphedCall(lambda *x: parts.part_b2(*x), list(t))
        return ent["b2"](*t, clone=False)

    return run_a, run_b1, run_b2


def meta_step(model, data, ridx=None, sync=False):
    """-> (predictions, losses, flat_grads): flat_grads["all"] is ONE [1, n_theta+n_psi+n_phi] buffer
    with the summed meta-gradients (not yet added to .grad; "theta"/"psi"/"phi" are its views, laid out by
    the loop's packs).  sync: all-reduce (SUM) it over the ranks in two buckets - phi, final after the
    dual pass, on a side stream while the 1-frame detector backward runs; then theta | psi (parallel.
    BucketedAllReduce); flat_grads["allreduce"] holds the CUDA events that time the collective."""
    loop = model._get_loop()
    ops = loop.ops
    if ops._clean:
        raise RuntimeError("the meta-training step needs the fp32-accurate GEMM mode (ITN_GEMM_PRECISION=tf32x3)")
    crit = model.criterion
    dev = ops.device
    frames = data["frames"].to(dev, non_blocking=True)
    masks = data["masks"]
    if not masks.is_cuda and dev.type == "cuda":
        from .episode import sample_masks_host
        masks = sample_masks_host(masks)              # only the h*w sampled mask pixels cross PCIe
    masks = masks.to(dev, non_blocking=True)
    E, S = frames.shape[:2]
    assert S == 5, "the meta-training step is defined on full 5-frame episodes"
    if model.mode == "train":
        model._new_dropout_seed(loop)
    C = loop.detector.class_embed.out_features
    NQ = detr_t.NQ
    tpk, ppk, fpk = loop.theta_pack, loop.psi_pack, loop.phi_pack
    kind = loop.kind
    if ridx is None:
        ridx = [random.randint(0, 4) for _ in range(E)]                     # reference :129, one draw per task
    idx = torch.tensor([e * S + int(r) for e, r in enumerate(ridx)]).to(dev)
    run_a, run_b1, run_b2 = _parts_runner(model, frames, masks, idx)

    # steps 1-3 (+ 1-frame pass): detector outputs with the fast weights ---------------------------
    a = run_a(frames, masks, idx)
    targets = _targets(data, E, S)
    keys5 = ("loss_ce", "class_error", "cardinality_error", "loss_bbox", "loss_giou")
    outs = {"pred_logits": a["post_logits"], "pred_boxes": a["post_boxes"]}
    sup_l, dlog, dbox = crit.loss_and_grad(outs, targets, background_c=0.1, groups=E, weights=LOSS_W)
    sup_host = sup_l.cpu()                                                        # [E,5]
    sup = {k: sup_host[:, i] for i, k in enumerate(keys5)}
    dact = None
    if kind == "A":
        # lowest-loss policy labels (reference models/interactron.py:105-118)
        f0 = {"pred_logits": a["post_logits"].view(E, S, NQ, C)[:, 0].contiguous(),
              "pred_boxes": a["post_boxes"].view(E, S, NQ, 4)[:, 0].contiguous()}
        t0 = [targets[e * S] for e in range(E)]
        gt = crit.group_losses(f0, t0, background_c=0.1, groups=E).cpu()
        rew = gt[:, 0] + 5 * gt[:, 4] + 2 * gt[:, 3]
        # the trie is process state fed by every episode of the (global) batch: replay all ranks' episodes
        # in global order, keep the labels of our own
        mine = [(data["initial_image_path"][e], [int(x) for x in data["actions"][e][:4]], float(rew[e]))
                for e in range(E)]
        rank = parallel.world()[0]
        best = [None] * E
        for r_, j, (iip, path, cost) in parallel.exchange_in_order(mine):
            store = model.path_storage.setdefault(iip, PathStorage())
            store.add_path(path, cost)
            if r_ == rank:
                best[j] = store.get_label(path)
        best = torch.tensor(best, dtype=torch.long).to(dev)                          # [E,4]
        # 16 logits per episode: CE over the 4 action heads and its gradient
        logp = F.log_softmax(a["actions"].view(E, 4, 4), dim=-1)
        loss_path = -logp.gather(-1, best[..., None]).squeeze(-1).mean(-1)            # [E]
        dact = ((logp.exp() - F.one_hot(best, 4).to(logp.dtype)) / 4.0).contiguous()
        sup["loss_path"] = loss_path.cpu()
        sup["policy_reward"] = rew
    t1 = [targets[e * S + int(r)] for e, r in enumerate(ridx)]
    o1 = {"pred_logits": a["post1_logits"], "pred_boxes": a["post1_boxes"]}
    det_l, dlog1, dbox1 = crit.loss_and_grad(o1, t1, background_c=0.1, groups=E, weights=LOSS_W)
    det_host = det_l.cpu()

    # steps 4-6: backward passes + the dual (second-order) pass -> flat meta-gradient -----------------
    n_t, n_p, n_f = tpk.numel, ppk.numel, fpk.numel
    G = ops.empty(1, n_t + n_p + n_f)              # callers keep views of it in .grad: a fresh buffer per step
    red = parallel.BucketedAllReduce() if sync else None
    gphi = run_b1(*([dlog, dbox] + ([dact] if dact is not None else [])))["Gphi"]
    G[:, n_t + n_p:].copy_(gphi)
    if red is not None:
        red.launch_async(G[:, n_t + n_p:])           # overlaps the detector pass below
    gtp = run_b2(dlog1, dbox1)["Gtp"]
    G[:, :n_t + n_p].copy_(gtp)
    flat = {"all": G, "theta": G[:, :n_t], "psi": G[:, n_t:n_t + n_p], "phi": G[:, n_t + n_p:]}
    if red is not None:
        red.finish(G[:, :n_t + n_p])
        flat["allreduce"] = red

    det = {k: det_host[:, i] for i, k in enumerate(keys5)}
    order = ("loss_ce", "class_error", "loss_bbox", "loss_giou", "cardinality_error")
    losses = {k.replace("loss", "loss_detector"): det[k].mean().to(dev) for k in order}
    sup_order = order + (("loss_path", "policy_reward") if kind == "A" else ())
    losses.update({k.replace("loss", "loss_supervisor"): sup[k].mean().to(dev) for k in sup_order})
    predictions = {"pred_logits": a["post1_logits"].view(E, 1, NQ, C).clone(),
                   "pred_boxes": a["post1_boxes"].view(E, 1, NQ, 4).clone()}
    return predictions, losses, flat


# fusion parameters the loss never reaches: their .grad stays None, as in the reference
UNUSED_PHI = {"A": ("model.pos_emb", "box_decoder.", "logit_decoder."),
              "B": ("pos_embed", "box_decoder.", "logit_decoder.", "action_decoder.")}


def accumulate_grads(model, flat):
    """Add the flat meta-gradients to `.grad` of the detector / fusion Parameters (sum semantics of
    repeated `.backward()` calls; a Parameter whose grad is None receives a view of the flat buffer)."""
    loop = model._get_loop()
    for key, pack, params in (("theta", loop.theta_pack, loop.theta_params), ("psi", loop.psi_pack, loop.psi_params),
                              ("phi", loop.phi_pack, loop.phi_params)):
        buf = flat[key]
        for name, p in zip(pack.names, params):
            if not p.requires_grad or (key == "phi" and name.startswith(UNUSED_PHI[loop.kind])):
                continue
            gv = pack.view(buf, name)[0]
            if p.grad is None:
                p.grad = gv
            else:
                p.grad.add_(gv)
