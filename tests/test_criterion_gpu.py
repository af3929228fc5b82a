"""C1: SetCriterion / HungarianMatcher on the device through the C ABI, against the
reference goldens and the CPU oracle (values 1e-5 relative - fp32 reductions in a different order;
assignments bit-exact)."""
import os

import pytest
import torch

from oracle import port
from oracle.cases import criterion_case

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-5


def _crit():
    from interactron_b200.criterion import HungarianMatcher, SetCriterion
    return SetCriterion(1235, HungarianMatcher(1.0, 5.0, 2.0)).cuda()


def _cuda_targets(targets):
    return [{k: v.cuda() for k, v in t.items()} for t in targets]


def _close(a, b, tol=TOL):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return (a - b).abs().max().item() <= tol * max(1.0, b.abs().max().item())


def test_criterion_matches_reference_goldens():
    gold = torch.load(os.path.join(GOLD, "criterion.pt"))
    crit = _crit()
    for (seed, frames), g in gold.items():
        logits, boxes, targets = criterion_case(seed, frames)
        out = {"pred_logits": logits.cuda(), "pred_boxes": boxes.cuda()}
        idx = crit.matcher(out, _cuda_targets(targets))
        for (i, j), (gi, gj) in zip(idx, g["indices"]):
            assert torch.equal(i, gi) and torch.equal(j, gj)
        losses = crit(out, _cuda_targets(targets), background_c=0.1)
        assert list(losses.keys()) == g["keys"]
        for k in g["keys"]:
            assert losses[k].dim() == 0 and _close(losses[k], g["losses"][k]), (seed, frames, k)
        vals, dl, db = crit.loss_and_grad(out, _cuda_targets(targets), 0.1)
        assert vals.shape == (1, 5) and dl.shape == logits.shape and db.shape == boxes.shape
        assert _close(db, g["dboxes"]) and _close(dl[..., ::97], g["dlogits_cols"])
        assert _close(dl.abs().sum(-1), g["dlogits_abs_rowsum"])


def test_criterion_groups_equal_single_calls_and_is_deterministic():
    """E episodes in one launch == E reference-style calls, each normalised by its own num_boxes."""
    crit = _crit()
    cases = [criterion_case(20 + e, 5) for e in range(3)]
    logits = torch.cat([c[0] for c in cases]).cuda()
    boxes = torch.cat([c[1] for c in cases]).cuda()
    targets = _cuda_targets([t for c in cases for t in c[2]])
    out = {"pred_logits": logits, "pred_boxes": boxes}
    vals, dl, db = crit.loss_and_grad(out, targets, 0.1, groups=3)
    vals2, dl2, db2 = crit.loss_and_grad(out, targets, 0.1, groups=3)
    assert torch.equal(vals, vals2) and torch.equal(dl, dl2) and torch.equal(db, db2)
    for e, (lg, bx, tg) in enumerate(cases):
        v1, dl1, db1 = crit.loss_and_grad({"pred_logits": lg.cuda(), "pred_boxes": bx.cuda()}, _cuda_targets(tg), 0.1)
        assert torch.equal(vals[e], v1[0])
        assert torch.equal(dl[5 * e:5 * e + 5], dl1) and torch.equal(db[5 * e:5 * e + 5], db1)


@pytest.mark.parametrize("bg", [0.1, 1.0])
def test_criterion_vs_oracle_with_autograd(bg):
    crit = _crit()
    logits, boxes, targets = criterion_case(33, 4)
    idx = port.hungarian_match(logits, boxes, targets)
    lr, br = logits.clone().requires_grad_(True), boxes.clone().requires_grad_(True)
    ref = port.set_criterion(lr, br, targets, idx, background_c=bg)
    w = (0.7, 1.3, 2.1)
    total = w[0] * ref["loss_ce"] + w[1] * ref["loss_bbox"] + w[2] * ref["loss_giou"]
    gl, gb = torch.autograd.grad(total, (lr, br))
    vals, dl, db = crit.loss_and_grad({"pred_logits": logits.cuda(), "pred_boxes": boxes.cuda()},
                                      _cuda_targets(targets), bg, weights=w)
    from interactron_b200.criterion import LOSS_KEYS
    for k, v in zip(LOSS_KEYS, vals[0]):
        assert _close(v, ref[k].detach()), k
    assert _close(dl, gl) and _close(db, gb)


def test_criterion_without_targets():
    """No ground truth at all: every row is "no object", num_boxes clamps to 1, class_error = 100."""
    crit = _crit()
    logits, boxes, _ = criterion_case(5, 2)
    empty = [{"labels": torch.zeros(0, dtype=torch.int64).cuda(), "boxes": torch.zeros(0, 4).cuda()} for _ in range(2)]
    out = crit({"pred_logits": logits.cuda(), "pred_boxes": boxes.cuda()}, empty)
    ref_ce = torch.nn.functional.cross_entropy(logits.flatten(0, 1), torch.full((100,), 1235))
    assert _close(out["loss_ce"], ref_ce)
    assert float(out["class_error"]) == 100.0 and float(out["loss_bbox"]) == 0.0 and float(out["loss_giou"]) == 0.0
    card = (logits.argmax(-1) != 1235).sum(1).float().mean()
    assert _close(out["cardinality_error"], card)


def test_models_carry_the_device_criterion():
    import interactron_b200 as ib
    from interactron_b200.criterion import SetCriterion
    m = ib.build_model(ib.default_config("interactron_random", weights="synthetic").MODEL)
    assert isinstance(m.criterion, SetCriterion)
    assert [k for k in m.state_dict() if k.startswith("criterion.")] == ["criterion.empty_weight"]
