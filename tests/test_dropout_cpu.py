"""train()-mode dropout (layers.DropCtx, csrc/itn_philox.cuh restated in oracle/philox.py).

The reference trainers run the models in train() mode (engine/interactron_trainer.py:73), where nn.Dropout
(p=0.1) follows every attention softmax, sub-layer output and FFN activation.  PyTorch's generator cannot be
reproduced, so the strong check injects OUR counter-based masks into the unmodified reference: F.dropout is
replaced, for the duration of one reference `predict()`, by a function that multiplies with the mask of
(seed, site = index of the dropout call in forward order, element position in OUR token-major layout).  With
identical masks the reference's autograd and this repo's hand-derived forward/backward (on the float64 torch
simulation of the kernels) must agree to round-off: that pins every dropout site, its position relative to the
residual adds, the 1/(1-p) scaling and the mask reuse in the backward pass.  Needs /root/reference."""
import inspect

import numpy as np
import pytest
import torch

from oracle import philox
from oracle import reference_harness as rh
from oracle.sim_ops import SimOps


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def test_mask_function_statistics():
    """Keep rate 1 - p, no correlation between sites / seeds / neighbouring rows."""
    m = philox.keep_mask(12345, 3, 2000, 361, 0.1)
    assert m.shape == (2000, 361) and abs(m.mean() - 0.9) < 2e-3
    assert abs(m.mean(0) - 0.9).max() < 0.04 and abs(m.mean(1) - 0.9).max() < 0.08
    other = philox.keep_mask(12345, 4, 2000, 361, 0.1)
    assert abs((m & other).mean() - 0.81) < 3e-3                      # independent sites
    assert abs((m & philox.keep_mask(12346, 3, 2000, 361, 0.1)).mean() - 0.81) < 3e-3
    assert abs((m[1:] & m[:-1]).mean() - 0.81) < 3e-3                 # neighbouring rows
    assert abs((m[:, 1:] & m[:, :-1]).mean() - 0.81) < 3e-3           # neighbouring columns (same Philox call)
    assert np.array_equal(philox.keep_mask(12345, 3, 7, 361, 0.1, row0=100), m[100:107])
    assert philox.keep_mask(1, 0, 64, 64, 0.0).all()


class _InjectedDropout:
    """Replaces torch.nn.functional.dropout while the reference runs: our mask function, our site numbering."""

    def __init__(self, seed):
        self.seed, self.site, self.calls = seed, 0, []

    def __call__(self, x, p=0.5, training=True, inplace=False):
        if not training or p == 0.0:
            return x
        site = self.site
        self.site += 1
        caller = inspect.stack()[1].function
        if caller == "multi_head_attention_forward" or x.dim() == 4:
            # attention probabilities [B*nh, Lq, Lk] (nn.MultiheadAttention) or [B, nh, T, T] (GPT): rows as ours
            rows = x.numel() // x.shape[-1]
            keep = torch.from_numpy(philox.keep_mask(self.seed, site, rows, x.shape[-1], p)).reshape(x.shape)
        elif x.shape[0] == 1:
            # GPT, batch-first [1, T, C]: rows = tokens
            keep = torch.from_numpy(philox.keep_mask(self.seed, site, x.shape[1], x.shape[2], p)).reshape(x.shape)
        else:
            # DETR / fusion-B layers, sequence-first [L, N, D]; ours is token-major: row = n * L + l
            L, N, D = x.shape
            keep = torch.from_numpy(philox.keep_mask(self.seed, site, N * L, D, p)).reshape(N, L, D).permute(1, 0, 2)
        self.calls.append((site, caller, tuple(x.shape)))
        return x * keep.to(x.dtype) * (1.0 / (1.0 - p))


@pytest.mark.skipif(not rh.reference_available(), reason="reference checkout not present")
@pytest.mark.parametrize("model_type", ["interactron_random", "interactron"])
def test_train_mode_predict_matches_reference_with_injected_masks(model_type, monkeypatch):
    import interactron_b200 as ib
    from interactron_b200.episode import InnerLoop
    from interactron_b200.synthetic import synthetic_episode
    cfg = ib.default_config(model_type, weights="synthetic")
    model = ib.build_model(cfg.MODEL).double()
    model._ops = SimOps(torch.float64)
    ref = rh.build_reference_model(model_type, model.state_dict()).double()
    ref.train()                                                     # what the reference trainers do
    data = synthetic_episode(0)
    data["frames"] = data["frames"].double()
    seed = 0x1234_5678_9ABC

    inj = _InjectedDropout(seed)
    passes = iter([InnerLoop.PASS_SITES["pre"], InnerLoop.PASS_SITES["post"]])
    det_fwd, fus_fwd = ref.detector.forward, ref.fusion.forward

    def det(*a, **k):
        inj.site = next(passes)
        return det_fwd(*a, **k)

    def fus(*a, **k):
        inj.site = InnerLoop.PASS_SITES["fusion"]
        return fus_fwd(*a, **k)

    monkeypatch.setattr(ref.detector, "forward", det)
    monkeypatch.setattr(ref.fusion, "forward", fus)
    monkeypatch.setattr(torch.nn.functional, "dropout", inj)
    torch.set_default_dtype(torch.float64)           # fusion B allocates with the default dtype
    try:
        tr = rh.reference_predict_with_trace(ref, data)
    finally:
        torch.set_default_dtype(torch.float32)
    monkeypatch.undo()
    # 6 encoder layers x 4 + 6 decoder layers x 6 dropout calls per detector pass
    n_det = 6 * 4 + 6 * 6
    n_fus = 4 * 6 if model_type == "interactron_random" else 1 + 4 * 3
    assert len(inj.calls) == 2 * n_det + n_fus

    model.train()
    loop = model._get_loop()
    loop.drop_seed = torch.tensor([seed], dtype=torch.int64)
    out = loop.adapt_detect(data["frames"], data["masks"], post_frames=(0,), want_trace=True, train=True)
    t = out["trace"]
    assert rel(t["pre_logits"][0], tr["pre"]["pred_logits"][0]) < 1e-10
    assert rel(out["learned_loss"][0], tr["learned_loss"]) < 1e-10
    g_ref = torch.cat([g.reshape(-1) for g in tr["grads"]])
    assert rel(t["g"][0], g_ref) < 1e-8                              # hand-derived backward with the same masks
    th_ref = torch.cat([p.reshape(-1) for p in tr["theta_prime"]])
    assert rel(t["theta_prime"][0], th_ref) < 1e-10
    assert rel(out["pred_logits"][0], tr["out"]["pred_logits"][0]) < 1e-8
    assert rel(out["pred_boxes"][0], tr["out"]["pred_boxes"][0]) < 1e-8
    # and dropout does something: eval() differs
    ev = loop.adapt_detect(data["frames"], data["masks"], post_frames=(0,))
    assert rel(ev["pred_logits"][0], tr["out"]["pred_logits"][0]) > 1e-3


@pytest.mark.skipif(not rh.reference_available(), reason="reference checkout not present")
@pytest.mark.parametrize("model_type", ["interactron_random"])  # (interactron passes too: 160 s; its dropout sites are pinned by the predict test)
def test_train_mode_forward_matches_reference_with_injected_masks(model_type, monkeypatch):
    """The meta-training step in train() mode, one episode: losses, predictions and every meta-gradient (second
    order through the dual-number pass, which must regenerate the pre-adapt and fusion masks) against the
    reference's `forward()` + double backward with OUR masks injected into its dropout calls (float64)."""
    import interactron_b200 as ib
    from interactron_b200.episode import InnerLoop
    from interactron_b200.synthetic import collate_episodes, synthetic_episode
    cfg = ib.default_config(model_type, weights="synthetic")
    model = ib.build_model(cfg.MODEL)
    ref = rh.build_reference_model(model_type, model.state_dict()).double()
    ref.train()
    model = model.double().train()
    sim = SimOps(torch.float64)
    model._ops = sim
    model.criterion._ops_override = sim
    model.criterion.matcher._ops_override = sim
    data = collate_episodes([synthetic_episode(0)])
    data["frames"] = data["frames"].double()
    data["boxes"] = [[b.double() for b in ep] for ep in data["boxes"]]
    ridx = [3]
    seed = 0x0FED_CBA9_8765

    inj = _InjectedDropout(seed)
    passes = iter([InnerLoop.PASS_SITES["pre"], InnerLoop.PASS_SITES["post"], InnerLoop.PASS_SITES["post1"]])
    det_fwd, fus_fwd = ref.detector.forward, ref.fusion.forward

    def det(*a, **k):
        inj.site = next(passes)
        return det_fwd(*a, **k)

    def fus(*a, **k):
        inj.site = InnerLoop.PASS_SITES["fusion"]
        return fus_fwd(*a, **k)

    monkeypatch.setattr(ref.detector, "forward", det)
    monkeypatch.setattr(ref.fusion, "forward", fus)
    monkeypatch.setattr(torch.nn.functional, "dropout", inj)
    torch.set_default_dtype(torch.float64)
    try:
        p_ref, l_ref, g_ref = rh.reference_forward_with_grads(ref, data, ridx)
    finally:
        torch.set_default_dtype(torch.float32)
    monkeypatch.undo()

    # the model draws its step seed from a host generator: make it return ours
    monkeypatch.setattr(type(model), "_new_dropout_seed",
                        lambda self, loop: setattr(loop, "drop_seed", torch.tensor([seed], dtype=torch.int64)))
    model.zero_grad(set_to_none=True)
    p, l = model(data, ridx=ridx)
    for k in l_ref:
        assert float(l[k]) == pytest.approx(float(l_ref[k]), rel=1e-6, abs=1e-9), k
    for k in ("pred_logits", "pred_boxes"):
        assert float((p[k] - p_ref[k]).norm() / p_ref[k].norm()) < 1e-8, k
    bad, n = [], 0
    for name, prm in model.named_parameters():
        gr = g_ref[name]
        if gr is None:
            assert prm.grad is None, name
            continue
        assert prm.grad is not None, name
        n += 1
        err = float((prm.grad - gr).norm())
        if err > 1e-7 * float(gr.norm()) + 1e-12:
            bad.append((name, err / max(float(gr.norm()), 1e-30)))
    assert not bad, sorted(bad, key=lambda t: -t[1])[:10]
    assert n > 250
