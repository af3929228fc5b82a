"""CPU (float64) checks of the dual-number (forward-over-reverse) execution that implements the
second-order half of the meta-training step (interactron_b200/dual.py) against the reference
modules' own double backward (`autograd.grad(..., create_graph=True)` + `.backward()`, reference
models/interactron.py:98-123).  Needs /root/reference."""
import os
import sys

import pytest
import torch

from oracle import reference_harness as rh
from oracle.sim_ops import SimOps

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_fusion_cpu import _build, _inputs, _mine_inputs  # noqa: E402

pytestmark = pytest.mark.skipif(not rh.reference_available(), reason="reference checkout not present")


def close(mine, ref, tol=1e-8, floor=1e-13):
    return float((mine.double() - ref.double()).norm()) <= tol * float(ref.double().norm()) + floor


@pytest.mark.parametrize("model_type,E", [("interactron_random", 2), ("interactron", 1)])
def test_fusion_second_order(model_type, E):
    """d/d(eps) grad_phi f(x + eps*v) + grad_phi <wa, actions>  ==  grad_phi [ <grad_x f, v> + <wa, actions> ]."""
    from interactron_b200 import fusion
    from interactron_b200.dual import Dual, DualOps, DualWeights
    from interactron_b200.layers import GradSink
    ref, W = _build(model_type)
    pack, flat = W.tuples[0][0], W.tuples[0][1]
    S = 5
    xs = _inputs(E, S, 21)
    vs = _inputs(E, S, 22)
    wa = torch.randn(E, 4, 4, generator=torch.Generator().manual_seed(5), dtype=torch.float64)
    named = dict(ref.named_parameters())
    live = [n for n in pack.names if named[n].requires_grad]
    torch.set_default_dtype(torch.float64)
    try:
        t = 0
        gx_ref = []
        for e in range(E):
            xin = [x[e:e + 1].clone().requires_grad_(True) for x in xs]
            o = ref(dict(zip(("embedded_memory_features", "box_features", "pred_logits", "pred_boxes"), xin)))
            gx = torch.autograd.grad(torch.norm(o["loss"]), xin, create_graph=True)
            t = t + sum((g * v[e:e + 1]).sum() for g, v in zip(gx, vs)) + (o["actions"] * wa[e]).sum()
            gx_ref.append([g.detach() for g in gx])
        g_ref = torch.autograd.grad(t, [named[n] for n in live], allow_unused=True)
    finally:
        torch.set_default_dtype(torch.float32)

    ops = DualOps(SimOps(torch.float64))
    mem, preds = _mine_inputs(*xs)
    mem_v, preds_v = _mine_inputs(*vs)
    DW = DualWeights((pack, flat, None, None, None))
    fwd = fusion.fusion_a_forward if model_type == "interactron" else fusion.fusion_b_forward
    bwd = fusion.fusion_a_backward if model_type == "interactron" else fusion.fusion_b_backward
    out, cache = fwd(ops, DW, Dual(mem, mem_v), Dual(preds, preds_v), E, S, 361)
    gphi = ops.zeros(1, pack.numel)
    dmemory, dpreds = bwd(ops, DW, cache, sink=GradSink(ops, pack, gphi, shared=True),
                          dactions=Dual(torch.zeros_like(wa), wa.clone()))
    # primal parts are the ordinary first-order results
    dp = dpreds.p.view(E, S, 50, 1496)
    for e in range(E):
        assert close(dp[e, ..., :256], gx_ref[e][1][0])
        assert close(dmemory.p[e], gx_ref[e][0][0].flatten(2).transpose(1, 2).reshape(S * 361, 256))
    bad = []
    n = 0
    for name, gr in zip(live, g_ref):
        mine = pack.view(gphi.t, name)[0]
        if gr is None:
            assert float(mine.abs().max()) == 0.0, name
            continue
        n += 1
        if not close(mine, gr):
            bad.append((name, float((mine - gr).norm() / gr.norm().clamp_min(1e-30))))
    assert not bad, bad[:10]
    assert n > 60


def test_detr_t_second_order():
    """theta carries the tangent v (per episode); the tangent of the in_proj (psi) gradient equals
    grad_psi <grad_theta f, v> of the reference detector's double backward."""
    from torch import nn
    from interactron_b200 import detr_t, modules as M, synthetic as S
    from interactron_b200.dual import Dual, DualOps, DualWeights
    from interactron_b200.layers import GradSink
    from interactron_b200.params import detector_packs
    cfg = rh.reference_config("interactron_random").MODEL

    class Holder(nn.Module):
        def __init__(self):
            super().__init__()
            self.detector = M.DetectorHolder(cfg.NUM_CLASSES)
            self.fusion = M.FusionBHolder(cfg)

    torch.manual_seed(0)
    holder = Holder()
    sd = S.synthetic_state_dict(holder)
    holder.load_state_dict(sd)
    ref = rh.build_reference_model("interactron_random", sd).double()
    holder = holder.double()
    data = S.synthetic_episode(3, frames=2)
    rh._load()
    from models.detr_models.util.misc import NestedTensor
    Fe = 2
    img, mask = data["frames"][0].double(), data["masks"][0]
    mask[1, :, 250:] = 1
    det = ref.detector
    theta_names = rh.reference_fast_weight_names(ref)
    named = dict(det.named_parameters())
    theta_ref = [named[n] for n in theta_names]
    out = det(NestedTensor(img, mask))
    gen = torch.Generator().manual_seed(1)
    ws = [torch.randn(out[k].shape, generator=gen).double()
          for k in ("pred_logits", "pred_boxes", "box_features", "embedded_memory_features")]
    wl, wb, wh, wm = ws
    loss = ((out["pred_logits"] * wl).sum() + (out["pred_boxes"] * wb).sum() + (out["box_features"] * wh).sum()
            + (out["embedded_memory_features"] * wm).sum())
    g = torch.autograd.grad(loss, theta_ref, create_graph=True)
    vgen = torch.Generator().manual_seed(2)
    v = [torch.randn(p.shape, generator=vgen).double() * 0.05 for p in theta_ref]
    t = sum((a * b).sum() for a, b in zip(g, v))
    tp, tparams, pp, pparams = detector_packs(holder.detector)
    gpsi_ref = torch.autograd.grad(t, [named[n] for n in pp.names])

    base = SimOps(torch.float64)
    ops = DualOps(base)
    tf = tp.pack(tparams, dtype=torch.float64).unsqueeze(0)
    pf = pp.pack(pparams, dtype=torch.float64).unsqueeze(0)
    vf = tp.pack(v, dtype=torch.float64).unsqueeze(0)
    DW = DualWeights((tp, tf, None, vf, None), (pp, pf, None, None, None))
    src = out["image_features"].detach()
    L = src.shape[2] * src.shape[3]
    src_tok = src.flatten(2).transpose(1, 2).reshape(1, Fe * L, 2048).contiguous()
    m19 = torch.nn.functional.interpolate(mask[None].float(), size=src.shape[-2:]).to(torch.bool)[0]
    pos = base.pos_embed_sine(m19).reshape(Fe * L, 256)
    kmask = m19.reshape(Fe, L).to(torch.uint8).contiguous()
    preds = ops.zeros(Fe * 50, 1496)
    o, cache = detr_t.detr_t_forward(ops, DW, src_tok, pos, kmask, 1, Fe, L, preds=preds)
    assert close(o["logits"].p.view(Fe, 50, -1), out["pred_logits"].detach())
    gpsi = ops.zeros(1, pp.numel)
    dpreds = torch.cat((wh, wl, wb), -1).reshape(Fe * 50, 1496).contiguous()
    dmem = wm.flatten(2).transpose(1, 2).reshape(1, Fe * L, 256).contiguous()
    detr_t.detr_t_backward(ops, DW, cache, GradSink(ops, pp, gpsi, shared=True),
                           dpreds=Dual(dpreds, torch.zeros_like(dpreds)), dmemory=Dual(dmem, torch.zeros_like(dmem)))
    bad = [(n, float((pp.view(gpsi.t, n)[0] - gr).norm() / gr.norm())) for n, gr in zip(pp.names, gpsi_ref)
           if not close(pp.view(gpsi.t, n)[0], gr)]
    assert not bad, bad
