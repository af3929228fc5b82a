"""GPU parity of the evaluator post-processing kernel (itn_detect_postprocess) and of the records built
from its output: against torch / torchvision (CPU) on the same inputs and against the golden records made
with the reference's own building blocks (tests/golden/evaluator_records.pt)."""
import os

import pytest
import torch

from oracle import cases, port
from test_evaluator_cpu import GOLD, same_records

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from interactron_b200.ops import CudaOps
    return CudaOps()


def test_detect_postprocess_kernel_vs_torchvision(ops):
    import torchvision
    seeds = list(range(8)) + [20, 21]
    batch = [cases.evaluator_case(s) for s in seeds]
    logits = torch.stack([b[0] for b in batch]).cuda()
    boxes = torch.stack([b[1] for b in batch]).cuda()
    # image 8: every query background; image 9: exact score ties between overlapping boxes
    logits[8, :, -1] += 50.0
    logits[9, 1] = logits[9, 0]
    boxes[9, 1] = boxes[9, 0]
    count, keep, score, cat, xyxy = (t.cpu() for t in ops.detect_postprocess(logits, boxes, 1235, 0.5))
    lg, bx = logits.cpu(), boxes.cpu()
    assert int(count[8]) == 0 and bool((keep[8] == -1).all())
    for i in range(len(seeds)):
        n = int(count[i])
        want_score, want_cat = lg[i].softmax(-1).max(-1)
        k = keep[i, :n].long()
        assert bool((keep[i, n:] == -1).all())
        assert torch.equal(cat[i, :n].long(), want_cat[k])
        assert torch.allclose(score[i, :n], want_score[k], rtol=2e-6, atol=0)
        assert torch.equal(xyxy[i, :n], port._xyxy(bx[i])[k])                       # bit-exact box conversion
        # NMS decisions: bit-exact against torchvision's CPU kernel fed the kernel's own scores
        all_score = torch.zeros(lg.shape[1])
        all_score[k] = score[i, :n]
        fg = torch.nonzero(want_cat != 1235)[:, 0]
        dev_score = want_score.clone()
        dev_score[k] = score[i, :n]
        ref_keep = fg[torchvision.ops.nms(port._xyxy(bx[i])[fg], dev_score[fg], 0.5)]
        if i != 9:      # (ties: torchvision's sort order between equal scores is its own business)
            assert torch.equal(k, ref_keep), i
        assert set(k.tolist()) <= set(fg.tolist())
    k9 = keep[9, :int(count[9])].tolist()
    assert not (0 in k9 and 1 in k9) and (1 not in k9)                              # the lower query index wins a tie


def test_records_match_reference_golden(ops):
    import interactron_b200 as ib
    from interactron_b200.evaluator import DetectionPostprocessor, image_detections
    gold = torch.load(GOLD)
    seeds = sorted(gold["cases"])
    batch = [cases.evaluator_case(s) for s in seeds]

    class Holder:                      # DetectionPostprocessor only needs `_get_ops()`
        def _get_ops(self):
            return ops

    preds = {"pred_logits": torch.stack([b[0] for b in batch]).cuda()[:, None],
             "pred_boxes": torch.stack([b[1] for b in batch]).cuda()[:, None]}
    post = DetectionPostprocessor(Holder())(preds)
    for i, s in enumerate(seeds):
        got = image_detections(post, i, batch[i][2], batch[i][3], f"img{s}", gold["class_ids"])
        same_records(got, gold["cases"][s], tol=2e-6)
