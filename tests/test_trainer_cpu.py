"""Trainer step on the flat buffers (trainer.MetaTrainerStep) vs the reference's own sequence
`clip_grad_norm_` + Adam(detector) + Adam(fusion) + zero_grad (engine/interactron_trainer.py:70-71,
106-110), float64 on the simulator backend: host logic, layout, aliasing, LR schedule."""
import copy
import math

import pytest
import torch

from oracle.sim_ops import SimOps


def _reference_iteration(ref, opt_d, opt_s, grads, clip):
    for (n, p) in ref.named_parameters():
        p.grad = None if grads[n] is None else grads[n].clone()
    torch.nn.utils.clip_grad_norm_(ref.parameters(), clip)
    opt_d.step()
    opt_s.step()
    opt_d.zero_grad()
    opt_s.zero_grad()


@pytest.mark.parametrize("model_type", ["interactron_random", "interactron"])
def test_trainer_step_matches_torch_adam(model_type):
    import interactron_b200 as ib
    from interactron_b200 import meta
    from interactron_b200.trainer import MetaTrainerStep
    torch.manual_seed(0)
    model = ib.build_model(ib.default_config(model_type, weights="synthetic").MODEL).eval().double()
    model._ops = SimOps(torch.float64)
    ref = copy.deepcopy(model)
    lr_d, lr_s, clip = 1e-3, 3e-3, 1.0
    opt_d = torch.optim.Adam(ref.detector.parameters(), lr=lr_d)
    opt_s = torch.optim.Adam(ref.fusion.parameters(), lr=lr_s)
    tr = MetaTrainerStep(model, lr_d, lr_s, clip)
    loop = model._get_loop()
    sizes = (loop.theta_pack.numel, loop.psi_pack.numel, loop.phi_pack.numel)
    before = {n: p.detach().clone() for n, p in model.named_parameters()}
    for it in range(3):
        G = torch.zeros(1, sum(sizes), dtype=torch.float64)       # padding between tensors stays zero, as in meta.py
        base = 0
        for pack in (loop.theta_pack, loop.psi_pack, loop.phi_pack):
            for nm in pack.names:
                pack.view(G[:, base:base + pack.numel], nm).normal_(0.0, 10.0 if it == 0 else 1e-4)   # clipped / not
            base += pack.numel
        flat = {"all": G, "theta": G[:, :sizes[0]], "psi": G[:, sizes[0]:sizes[0] + sizes[1]], "phi": G[:, sizes[0] + sizes[1]:]}
        model.last_meta_grads = flat
        meta.accumulate_grads(model, flat)
        grads = {n: (None if p.grad is None else p.grad.clone()) for n, p in model.named_parameters()}
        assert sum(g is None for g in grads.values()) > 50           # frozen backbone + unused fusion heads
        want_norm = torch.sqrt(sum((g ** 2).sum() for g in grads.values() if g is not None))
        out = tr.step()
        assert float(out["grad_norm"]) == pytest.approx(float(want_norm), rel=1e-12)
        _reference_iteration(ref, opt_d, opt_s, grads, clip)
        assert all(p.grad is None for p in model.parameters())
        for (n, p), (_, q) in zip(model.named_parameters(), ref.named_parameters()):
            assert torch.allclose(p, q, rtol=0, atol=1e-13), (it, n)
    moved = sum(int(not torch.equal(p, before[n])) for n, p in model.named_parameters())
    assert moved > 250
    # the Parameters alias the flat buffers, and the W^T twins follow the update
    name = "transformer.encoder.layers.0.linear1.weight"
    w = dict(model.detector.named_parameters())[name]
    assert w.data_ptr() == loop.theta_pack.view(loop.theta, name).data_ptr()
    assert torch.equal(loop.theta_pack.view_t(loop.theta_t, name)[0], w.t())
    assert torch.equal(model.state_dict()["detector." + name], w)
    # gradients that do not alias the flat buffer (set by hand) take the gather path
    for p in model.parameters():
        p.grad = None
    model.last_meta_grads = None
    grads = {}
    for n, p in model.named_parameters():
        if n.startswith("fusion.loss_decoder") or n.endswith("norm1.weight"):
            p.grad = torch.randn_like(p)
        grads[n] = None if p.grad is None else p.grad.clone()
    tr.step()
    _reference_iteration(ref, opt_d, opt_s, grads, clip)
    for (n, p), (_, q) in zip(model.named_parameters(), ref.named_parameters()):
        assert torch.allclose(p, q, rtol=0, atol=1e-13), n


def test_supervisor_lr_schedule():
    """Warm-up / cosine decay by frames seen, applied to the supervisor LR only (reference :113-124)."""
    import interactron_b200 as ib
    from interactron_b200.trainer import MetaTrainerStep
    model = ib.build_model(ib.default_config("interactron_random", weights="synthetic").MODEL).eval().double()
    model._ops = SimOps(torch.float64)
    tr = MetaTrainerStep(model, 1e-5, 1e-4, 1.0, lr_decay=True, warmup_tokens=160, final_tokens=800)
    tokens, lrs = 0, []
    for it in range(12):
        lrs.append(tr.step(n_frames=80)["lr"])
    want, lr = [], 1e-4
    for it in range(12):
        want.append(lr)
        tokens += 80
        if tokens < 160:
            mult = tokens / 160
        else:
            mult = max(0.1, 0.5 * (1.0 + math.cos(math.pi * (tokens - 160) / (800 - 160))))
        lr = 1e-4 * mult
    assert lrs == pytest.approx(want, rel=1e-12)
    assert tr.detector_lr == 1e-5


def _reference_record(saved, sd, w):
    """engine/interactron_trainer.py:48-57 verbatim semantics."""
    if saved is None:
        return {k: w * v for k, v in sd.items()}
    for k, v in sd.items():
        saved[k] += w * v
    return saved


@pytest.mark.parametrize("model_type", ["interactron_random", "interactron"])
def test_windowed_checkpoint_average(model_type):
    """trainer.CheckpointAverager on the flat buffers == the reference's record_checkpoint / save_checkpoint
    dict arithmetic over the last SAVE_WINDOW epochs (reference key layout and order)."""
    import interactron_b200 as ib
    from interactron_b200 import meta
    from interactron_b200.trainer import CheckpointAverager, MetaTrainerStep
    torch.manual_seed(1)
    model = ib.build_model(ib.default_config(model_type, weights="synthetic").MODEL).eval().double()
    model._ops = SimOps(torch.float64)
    tr = MetaTrainerStep(model, 1e-3, 3e-3, 1.0)
    avg = CheckpointAverager(model)
    assert list(avg.state_dict().keys()) == list(model.state_dict().keys())     # nothing recorded: plain state_dict
    loop = model._get_loop()
    sizes = (loop.theta_pack.numel, loop.psi_pack.numel, loop.phi_pack.numel)
    window, saved = 3, None
    for epoch in range(window):
        G = torch.randn(1, sum(sizes), dtype=torch.float64)
        flat = {"all": G, "theta": G[:, :sizes[0]], "psi": G[:, sizes[0]:sizes[0] + sizes[1]], "phi": G[:, sizes[0] + sizes[1]:]}
        model.last_meta_grads = flat
        meta.accumulate_grads(model, flat)
        tr.step()                                                                # weights move between the records
        saved = _reference_record(saved, {k: v.clone() for k, v in model.state_dict().items()}, 1.0 / window)
        avg.record_checkpoint(1.0 / window)
    out = avg.state_dict()
    assert list(out.keys()) == list(saved.keys()) == list(model.state_dict().keys())
    n_flat = 0
    for k in saved:
        assert out[k].shape == saved[k].shape, k
        assert torch.allclose(out[k].double(), saved[k].double(), rtol=0, atol=1e-14), k
        n_flat += int(out[k].data_ptr() != saved[k].data_ptr())
    moved = [k for k in saved if not torch.equal(saved[k], model.state_dict()[k].to(saved[k].dtype))]
    assert len(moved) > 250                                                       # it is an average, not the last snapshot


def test_optimizer_state_roundtrip_and_step_count_guard():
    """state_dict()/load_state_dict() resume the fused step bit-exactly; a parameter that skipped steps and then
    receives a gradient (torch.optim.Adam would use its own step count) is refused instead of silently diverging."""
    import interactron_b200 as ib
    from interactron_b200.trainer import MetaTrainerStep
    torch.manual_seed(1)

    def make():
        m = ib.build_model(ib.default_config("interactron_random", weights="synthetic").MODEL).eval().double()
        m._ops = SimOps(torch.float64)
        return m, MetaTrainerStep(m, 1e-3, 3e-3, 1.0)

    def set_grads(m, seed, only=None):
        g = torch.Generator().manual_seed(seed)
        m.last_meta_grads = None
        for n, p in m.named_parameters():
            take = n.startswith("detector.transformer.decoder") if only is None else only(n)
            p.grad = torch.randn(p.shape, generator=g, dtype=p.dtype) if take else None

    a, ta = make()
    for it in range(2):
        set_grads(a, it)
        ta.step()
    sd, weights = ta.state_dict(), copy.deepcopy(a.state_dict())
    set_grads(a, 7)
    ta.step()
    b, tb = make()
    b.load_state_dict(weights)
    tb.load_state_dict(sd)
    assert tb.t == 2
    set_grads(b, 7)
    tb.step()
    for (n, p), (_, q) in zip(a.named_parameters(), b.named_parameters()):
        assert torch.equal(p, q), n
    # a parameter group that never had a gradient gets one at step 4: refused
    set_grads(a, 8, only=lambda n: n.startswith("detector.transformer.encoder"))
    with pytest.raises(RuntimeError, match="skipped earlier steps"):
        ta.step()
