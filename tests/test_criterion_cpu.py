"""oracle/port.py SetCriterion + HungarianMatcher restatement pinned to the reference: against the
committed goldens (tests/golden/criterion.pt, made by tools/make_golden.py from the unmodified
reference) and, where /root/reference exists, against the reference run live."""
import os

import pytest
import torch

from oracle import port
from oracle.cases import criterion_case

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _total(out):
    return out["loss_ce"] + 5 * out["loss_giou"] + 2 * out["loss_bbox"]      # models/interactron.py:121-122


def test_port_criterion_matches_goldens():
    gold = torch.load(os.path.join(GOLD, "criterion.pt"))
    assert len(gold) == 6
    for (seed, frames), g in gold.items():
        logits, boxes, targets = criterion_case(seed, frames)
        idx = port.hungarian_match(logits, boxes, targets)
        for (i, j), (gi, gj) in zip(idx, g["indices"]):
            assert torch.equal(i, gi) and torch.equal(j, gj)                 # bit-exact assignments
        logits.requires_grad_(True)
        boxes.requires_grad_(True)
        out = port.set_criterion(logits, boxes, targets, idx)
        assert list(out.keys()) == g["keys"]
        for k in g["keys"]:
            assert float(out[k].detach()) == pytest.approx(float(g["losses"][k]), rel=1e-6, abs=1e-7), k
        dl, db = torch.autograd.grad(_total(out), (logits, boxes))
        assert (db - g["dboxes"]).abs().max() < 1e-7
        assert (dl[..., ::97] - g["dlogits_cols"]).abs().max() < 1e-7
        assert (dl.abs().sum(-1) - g["dlogits_abs_rowsum"]).abs().max() < 1e-6


@pytest.mark.skipif(not os.path.isdir("/root/reference/models"), reason="reference tree not present")
def test_port_criterion_matches_live_reference():
    from oracle import reference_harness as rh
    rh._load()
    from models.detr_models.detr import SetCriterion
    from models.detr_models.matcher import HungarianMatcher
    crit = SetCriterion(1235, matcher=HungarianMatcher(1, 5, 2), weight_dict={}, eos_coef=0.1,
                        losses=["labels", "boxes", "cardinality"])
    logits, boxes, targets = criterion_case(11, 3)
    for bg in (0.1, 0.5):
        ref = crit({"pred_logits": logits, "pred_boxes": boxes}, targets, background_c=bg)
        idx = port.hungarian_match(logits, boxes, targets)
        out = port.set_criterion(logits, boxes, targets, idx, background_c=bg)
        for k, v in ref.items():
            assert float(out[k]) == pytest.approx(float(v), rel=1e-6), k
