"""Evaluator post-processing, host side: the oracle restatement and the product's record
builder against the golden records produced with the reference's own building blocks
(tools/make_golden_eval.py), and both matcher restatements against the reference function itself."""
import os

import numpy as np
import pytest
import torch

from oracle import cases, port
from oracle import reference_harness as rh

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "evaluator_records.pt")


def same_records(got, want, tol=0.0):
    assert len(got) == len(want), (len(got), len(want))
    for g, w in zip(got, want):
        assert list(g.keys()) == list(w.keys())
        for k in ("type", "category_match", "pred_cat", "img"):
            assert g[k] == w[k], (k, g, w)
        for k in ("iou", "pred_score", "area"):
            assert g[k] == pytest.approx(w[k], rel=tol, abs=tol), (k, g, w)
        assert g["box"] == pytest.approx(w["box"], rel=tol, abs=tol)


def simulated_post(logits, boxes, background=1235):
    """What itn_detect_postprocess returns, computed with torch / torchvision on the CPU."""
    import torchvision
    xyxy = port._xyxy(boxes)
    score, cat = logits.softmax(-1).max(-1)
    idx = torch.nonzero(cat != background)[:, 0]
    kept = idx[torchvision.ops.nms(xyxy[idx], score[idx], 0.5)]
    Q = logits.shape[0]
    post = {"count": np.array([len(kept)]), "keep_idx": np.full((1, Q), -1), "score": np.zeros((1, Q), np.float32),
            "cat": np.zeros((1, Q), np.int32), "xyxy": np.zeros((1, Q, 4), np.float32)}
    post["keep_idx"][0, :len(kept)] = kept.numpy()
    post["score"][0, :len(kept)] = score[kept].numpy()
    post["cat"][0, :len(kept)] = cat[kept].numpy()
    post["xyxy"][0, :len(kept)] = xyxy[kept].numpy()
    return post


def test_oracle_and_host_records_match_golden():
    from interactron_b200.evaluator import image_detections
    gold = torch.load(GOLD)
    kinds = set()
    for seed, want in gold["cases"].items():
        logits, boxes, gt_boxes, gt_cats = cases.evaluator_case(seed)
        same_records(port.evaluator_records(logits, boxes, gt_boxes, gt_cats, f"img{seed}", gold["class_ids"]), want)
        got = image_detections(simulated_post(logits, boxes), 0, gt_boxes, gt_cats, f"img{seed}", gold["class_ids"])
        same_records(got, want, tol=1e-7)
        kinds |= {(r["type"], r["category_match"]) for r in want}
    assert kinds == {("tp", True), ("fp", True), ("fn", False), ("fp", False)}


@pytest.mark.skipif(not rh.reference_available(), reason="reference checkout not present")
def test_matchers_equal_the_reference_function():
    from interactron_b200.evaluator import match_predictions
    rh._load()
    from utils.detection_utils import match_predictions_to_detections as ref_match
    gen = torch.Generator().manual_seed(5)
    for trial in range(200):
        p, g = int(torch.randint(1, 7, (1,), generator=gen)), int(torch.randint(1, 6, (1,), generator=gen))
        ious = torch.rand(p, g, generator=gen)
        ious = ious * (torch.rand(p, g, generator=gen) < 0.5)             # many exact zeros (non-overlapping boxes)
        if trial % 5 == 0:
            ious = (ious * 4).round() / 4                                  # exact ties
        rb, ri = ref_match(ious.clone())
        ob, oi = port.match_predictions_to_detections(ious.clone())
        mb, mi = match_predictions(ious.numpy())
        assert torch.equal(rb, ob) and torch.equal(ri, oi), trial
        assert np.array_equal(rb.numpy(), mb) and np.array_equal(ri.numpy(), mi), trial
