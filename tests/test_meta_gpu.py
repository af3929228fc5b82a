"""GPU parity of the meta-training step `forward(data)` (second-order supervisor + first-order
detector meta-gradients accumulated on .grad) against the goldens of the unmodified reference
(tests/golden/<model>_forward.pt from tools/make_golden_meta.py; D1 mode, eval()).

Predictions and losses: 1e-3 relative (the north star's bar).  Meta-gradients: the fixtures hold the
reference run in fp32 AND in fp64.  These gradients are ill-conditioned (a second derivative through
12 post-norm layers): the fp32 reference itself sits 2e-4 (`interactron_random`) / 2e-3 (`interactron`) away
from its own fp64 run in relative L2 over all parameters.  `forward()` runs its GEMMs with split accumulators
(ITN_PREC_TF32X3_SPLIT, models.meta_split_acc) and the two-term round-toward-zero compensation: measured on B200
(tools/meta_parity.py, profiles/README.md) 6.7e-4 / 2.3e-3 over all parameters, per group 4.4e-4 ... 7.5e-4 and
1.6e-3 ... 3.0e-3 (fp32 reference: 0.9e-4 ... 2.2e-4 and 1.1e-3 ... 2.9e-3).  History: round 1 9.4e-4 / 3.8e-3 with bounds
max(3e-3, 8 x gap); the same step on the fp32 FMA GEMM (ITN_FORCE_SIMT=1): 1.2e-4, so the derivation and every
non-GEMM kernel are exact to fp32.  Bounds, against the fp64 reference (the ones VERDICT round 1 asked for):
  per group (theta, psi, phi):   max(1e-3, 3 x fp32-reference gap)
  per tensor (strided sample):   max(3e-2, 10 x fp32-reference gap); exact zeros stay (near) zero.
The default tf32x3 mode (`meta_split_acc = False`) is checked against max(1.5e-3, 3 x gap) (measured 1.07e-3)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _data(gold):
    from interactron_b200.synthetic import collate_episodes, synthetic_episode
    return collate_episodes([synthetic_episode(e) for e in gold["episodes"]])


def _check_round(model, data, gold, r, floor=1e-3):
    g32, g64 = gold["fp32"][r], gold["fp64"][r]
    model.zero_grad(set_to_none=True)
    p, l = model(data, ridx=list(gold["ridx"]))
    torch.cuda.synchronize()
    assert list(l.keys()) == list(g32["losses"].keys())
    for k, v in g64["losses"].items():
        assert float(l[k]) == pytest.approx(float(v), rel=1e-3, abs=1e-5), k
    for k in ("pred_logits", "pred_boxes"):
        assert p[k].shape == g32["predictions"][k].shape
        assert rel(p[k], g32["predictions"][k]) < 1e-3, k
    stride_of = lambda n: max(1, n // 256) | 1        # tools/make_golden_meta.py
    groups, per = {}, []
    top = max(float(e["norm"]) for e in g64["grads"].values() if e is not None)
    for name, prm in model.named_parameters():
        e64, e32 = g64["grads"][name], g32["grads"][name]
        if e64 is None:
            assert prm.grad is None, name
            continue
        assert prm.grad is not None, name
        mine = prm.grad.detach().double().cpu().reshape(-1)
        s64, s32 = e64["sample"].double(), e32["sample"].double()
        ms = mine[::stride_of(mine.numel())]
        key = "phi" if name.startswith("fusion.") else ("psi" if "in_proj" in name else "theta")
        a = groups.setdefault(key, [0.0, 0.0, 0.0])
        a[0] += float((ms - s64).pow(2).sum())
        a[1] += float((s32 - s64).pow(2).sum())
        a[2] += float(s64.pow(2).sum())
        if float(e64["norm"]) < 1e-6 * top:        # (near-)zero gradients, e.g. key biases: stay small
            assert float(mine.norm()) < 1e-4 * top, name
            continue
        gap = float((s32 - s64).norm() / s64.norm().clamp_min(1e-30))
        err = float((ms - s64).norm() / s64.norm().clamp_min(1e-30))
        per.append((err / max(3e-2, 10 * gap), err, gap, name))
    per.sort(reverse=True)
    print("worst tensors (err/bound, err vs fp64, fp32-reference gap, name):")
    for w in per[:6]:
        print("   ", w)
    for key, a in groups.items():
        mine_e, ref_e = (a[0] / a[2]) ** 0.5, (a[1] / a[2]) ** 0.5
        print(f"group {key}: ours vs fp64 {mine_e:.3e}, fp32 reference vs fp64 {ref_e:.3e}")
        assert mine_e < max(floor, 3 * ref_e), (key, mine_e, ref_e)
    assert per[0][0] < 1.0, per[:5]


@pytest.mark.parametrize("name", ["interactron_random", "interactron"])
def test_forward_matches_reference_goldens(name):
    import interactron_b200 as ib
    gold = torch.load(os.path.join(GOLD, f"{name}_forward.pt"))
    model = ib.build_model(ib.default_config(name, weights="synthetic").MODEL).cuda().eval()
    data = _data(gold)
    for r in range(len(gold["fp32"])):
        _check_round(model, data, gold, r)
    # parameters are untouched; grads accumulate (sum) over calls like repeated .backward()
    g1 = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
    model(data, ridx=list(gold["ridx"]))
    some = [n for n in g1 if "in_proj_weight" in n][:2] + [n for n in g1 if n.startswith("fusion.")][:2]
    if name == "interactron_random":      # (fusion A's second call sees a populated path storage)
        for n in some:
            assert rel(dict(model.named_parameters())[n].grad, 2 * g1[n]) < 1e-5, n


def test_forward_default_gemm_mode_is_close_too():
    """The same step with the default tf32x3 GEMMs of predict() (no split accumulators): 1.07e-3 on the worst group."""
    import interactron_b200 as ib
    gold = torch.load(os.path.join(GOLD, "interactron_random_forward.pt"))
    model = ib.build_model(ib.default_config("interactron_random", weights="synthetic").MODEL).cuda().eval()
    model.meta_split_acc = False
    _check_round(model, _data(gold), gold, 0, floor=1.5e-3)


def test_forward_batch_equals_singles():
    """Meta-gradients are summed over the episodes of the batch: E=2 in one call == two E=1 calls."""
    import interactron_b200 as ib
    from interactron_b200.synthetic import collate_episodes, synthetic_episode
    model = ib.build_model(ib.default_config("interactron_random", weights="synthetic").MODEL).cuda().eval()
    eps = [synthetic_episode(2), synthetic_episode(3)]
    model.zero_grad(set_to_none=True)
    model(collate_episodes(eps), ridx=[0, 4])
    both = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
    model.zero_grad(set_to_none=True)
    model(collate_episodes(eps[:1]), ridx=[0])
    model(collate_episodes(eps[1:]), ridx=[4])
    top = max(float(g.norm()) for g in both.values())
    for n, p in model.named_parameters():
        if p.grad is not None and float(both[n].norm()) > 1e-5 * top:     # skip the mathematically-zero ones
            assert rel(p.grad, both[n]) < 2e-3, n        # different batch shapes -> different GEMM tilings
