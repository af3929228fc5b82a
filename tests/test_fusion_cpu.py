"""CPU (float64) check of the fusion networks' forward + hand-derived data-gradient backward
(interactron_b200.fusion) against the reference modules' autograd.  Needs /root/reference."""
import pytest
import torch

from oracle import reference_harness as rh
from oracle.sim_ops import SimOps

pytestmark = pytest.mark.skipif(not rh.reference_available(), reason="reference checkout not present")


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def _build(model_type):
    from torch import nn
    from interactron_b200 import modules as M, synthetic as S
    from interactron_b200.params import ParamPack, Weights
    cfg = rh.reference_config(model_type).MODEL
    rh._load()
    if model_type == "interactron":
        from models.transformer import Transformer
        holder = M.FusionAHolder(cfg)
    else:
        from models.new_transformer import Transformer
        holder = M.FusionBHolder(cfg)

    class Wrap(nn.Module):
        def __init__(self):
            super().__init__()
            self.fusion = holder

    sd = S.synthetic_state_dict(Wrap())
    fsd = {k[len("fusion."):]: v for k, v in sd.items()}
    holder.load_state_dict(fsd)
    ref = Transformer(cfg)
    ref.load_state_dict(fsd)
    ref = ref.double().eval()
    holder = holder.double()
    items = [(n, p) for n, p in holder.named_parameters()]
    pack = ParamPack(items)
    flat = pack.pack([p for _, p in items], dtype=torch.float64).unsqueeze(0)
    return ref, Weights((pack, flat, flat))


def _inputs(E, S, seed):
    gen = torch.Generator().manual_seed(seed)
    mem = torch.randn(E, S, 256, 19, 19, generator=gen, dtype=torch.float64)
    bf = torch.randn(E, S, 50, 256, generator=gen, dtype=torch.float64)
    lg = torch.randn(E, S, 50, 1236, generator=gen, dtype=torch.float64)
    bx = torch.rand(E, S, 50, 4, generator=gen, dtype=torch.float64)
    return mem, bf, lg, bx


def _ref_run(ref, mem, bf, lg, bx):
    # fusion B allocates its padded buffers with the default dtype (reference new_transformer.py:41-44)
    torch.set_default_dtype(torch.float64)
    try:
        return _ref_run_inner(ref, mem, bf, lg, bx)
    finally:
        torch.set_default_dtype(torch.float32)


def _ref_run_inner(ref, mem, bf, lg, bx):
    outs = []
    for e in range(mem.shape[0]):
        x = {"embedded_memory_features": mem[e:e + 1].clone().requires_grad_(True),
             "box_features": bf[e:e + 1].clone().requires_grad_(True),
             "pred_logits": lg[e:e + 1].clone().requires_grad_(True),
             "pred_boxes": bx[e:e + 1].clone().requires_grad_(True)}
        o = ref(x)
        ll = torch.norm(o["loss"])
        g = torch.autograd.grad(ll, list(x.values()))
        outs.append((o, ll, g))
    return outs


def _mine_inputs(mem, bf, lg, bx):
    E, S = mem.shape[:2]
    memory_r = mem.flatten(3).transpose(2, 3).reshape(E, S * 361, 256).contiguous()
    preds = torch.cat((bf, lg, bx), -1).reshape(E * S * 50, 1496).contiguous()
    return memory_r, preds


@pytest.mark.parametrize("model_type,E", [("interactron_random", 2), ("interactron", 1)])
def test_fusion_forward_backward(model_type, E):
    from interactron_b200 import fusion
    ref, W = _build(model_type)
    S = 5
    mem, bf, lg, bx = _inputs(E, S, 7)
    refs = _ref_run(ref, mem, bf, lg, bx)
    ops = SimOps(torch.float64)
    memory_r, preds = _mine_inputs(mem, bf, lg, bx)
    fwd = fusion.fusion_a_forward if model_type == "interactron" else fusion.fusion_b_forward
    bwd = fusion.fusion_a_backward if model_type == "interactron" else fusion.fusion_b_backward
    out, cache = fwd(ops, W, memory_r, preds, E, S, 361)
    dmemory, dpreds = bwd(ops, W, cache)
    dp = dpreds.view(E, S, 50, 1496)
    for e, (o, ll, g) in enumerate(refs):
        assert rel(out["loss_vec"][e], o["loss"].reshape(-1)) < 1e-10
        assert rel(out["learned_loss"][e], ll) < 1e-10
        assert rel(out["actions"][e], o["actions"]) < 1e-10
        gm = g[0][0].flatten(2).transpose(1, 2).reshape(S * 361, 256)
        assert rel(dmemory[e], gm) < 1e-9
        assert rel(dp[e, ..., :256], g[1][0]) < 1e-9
        assert rel(dp[e, ..., 256:1492], g[2][0]) < 1e-9
        assert rel(dp[e, ..., 1492:], g[3][0]) < 1e-9


def test_fusion_a_partial_sequence_and_aux_heads():
    """Policy rollout feeds 1-4 frames (T = 411*s + 5); the multi-frame baseline reads the
    box/logit decoders."""
    from interactron_b200 import fusion
    ref, W = _build("interactron")
    S = 2
    mem, bf, lg, bx = _inputs(1, S, 9)
    with torch.no_grad():
        o = ref({"embedded_memory_features": mem, "box_features": bf, "pred_logits": lg, "pred_boxes": bx})
    ops = SimOps(torch.float64)
    memory_r, preds = _mine_inputs(mem, bf, lg, bx)
    out, _ = fusion.fusion_a_forward(ops, W, memory_r, preds, 1, S, 361, need_cache=False, want_aux_heads=True)
    assert rel(out["actions"][0], o["actions"]) < 1e-10
    assert rel(out["pred_boxes"].view(S, 50, 4), o["pred_boxes"]) < 1e-10
    assert rel(out["pred_logits"].view(S, 50, 1236), o["pred_logits"]) < 1e-10


@pytest.mark.parametrize("model_type,E", [("interactron_random", 2), ("interactron", 1)])
def test_fusion_parameter_gradients(model_type, E):
    """phi gradients of sum_e (learned_loss_e + <wa_e, actions_e>) summed over the episodes, as
    `.backward()` accumulates them in the meta-training step (reference models/interactron.py:118-123)."""
    from interactron_b200 import fusion
    from interactron_b200.layers import GradSink
    ref, W = _build(model_type)
    pack = W.tuples[0][0]
    S = 5
    mem, bf, lg, bx = _inputs(E, S, 11)
    wa = torch.randn(E, 4, 4, generator=torch.Generator().manual_seed(5), dtype=torch.float64)
    torch.set_default_dtype(torch.float64)
    try:
        total = 0
        for e in range(E):
            o = ref({"embedded_memory_features": mem[e:e + 1], "box_features": bf[e:e + 1],
                     "pred_logits": lg[e:e + 1], "pred_boxes": bx[e:e + 1]})
            total = total + torch.norm(o["loss"]) + (o["actions"] * wa[e]).sum()
    finally:
        torch.set_default_dtype(torch.float32)
    named = dict(ref.named_parameters())
    live = [n for n in pack.names if named[n].requires_grad]
    g_ref = torch.autograd.grad(total, [named[n] for n in live], allow_unused=True)
    ops = SimOps(torch.float64)
    memory_r, preds = _mine_inputs(mem, bf, lg, bx)
    fwd = fusion.fusion_a_forward if model_type == "interactron" else fusion.fusion_b_forward
    bwd = fusion.fusion_a_backward if model_type == "interactron" else fusion.fusion_b_backward
    out, cache = fwd(ops, W, memory_r, preds, E, S, 361)
    gphi = ops.zeros(1, pack.numel)
    bwd(ops, W, cache, sink=GradSink(ops, pack, gphi, shared=True), dactions=wa.clone())
    n_checked = 0
    for name, gr in zip(live, g_ref):
        mine = pack.view(gphi, name)[0]
        if gr is None:
            assert float(mine.abs().max()) == 0.0, name
            continue
        # key biases have an exactly-zero gradient (softmax shift invariance): absolute floor
        assert float((mine - gr).norm()) <= 1e-9 * float(gr.norm()) + 1e-13, name
        n_checked += 1
    assert n_checked > 60
