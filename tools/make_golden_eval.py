"""Golden TP / FP / FN records for the evaluator post-processing.  The reference's
evaluate() (engine/random_policy_evaluator.py:37-175) is one monolithic method bound to a dataset, so its
per-image body is restated in oracle/port.evaluator_records; this script pins that restatement by running it
with the REFERENCE'S OWN building blocks - utils.detection_utils.match_predictions_to_detections,
detr_models.util.box_ops.box_cxcywh_to_xyxy, utils.constants.THOR_CLASS_IDS - on oracle/cases.evaluator_case.
-> tests/golden/evaluator_records.pt        (python tools/make_golden_eval.py, build container only)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import cases, port  # noqa: E402
from oracle import reference_harness as rh  # noqa: E402

if __name__ == "__main__":
    rh._load()
    from models.detr_models.util.box_ops import box_cxcywh_to_xyxy
    from utils.constants import THOR_CLASS_IDS
    from utils.detection_utils import match_predictions_to_detections
    real_xyxy = port._xyxy
    port._xyxy = box_cxcywh_to_xyxy                       # the reference's conversion inside the restated body
    try:
        gold = {"class_ids": list(THOR_CLASS_IDS), "cases": {}}
        for seed in range(8):
            logits, boxes, gt_boxes, gt_cats = cases.evaluator_case(seed)
            recs = port.evaluator_records(logits, boxes, gt_boxes, gt_cats, f"img{seed}", THOR_CLASS_IDS,
                                          match_fn=match_predictions_to_detections)
            gold["cases"][seed] = recs
            kinds = [r["type"] for r in recs]
            print(seed, "gt", len(gt_cats), {k: kinds.count(k) for k in ("tp", "fp", "fn")})
    finally:
        port._xyxy = real_xyxy
    torch.save(gold, os.path.join(ROOT, "tests", "golden", "evaluator_records.pt"))
