"""tests/golden/interactron_random_predict_unfrozen.pt: predict() of the UNMODIFIED reference WITHOUT decision
D1, i.e. with the ResNet layer2-4 convolutions among the fast weights (199 tensors instead of 157; reference
models/detr_models/backbone.py:61-63, utils/meta_utils.py:9-15).  This repo adapts the 157 transformer / head
tensors only (D1), so this fixture is the yardstick of that gap, not a parity target:
tests/test_predict_gpu.py carries it as a strict xfail.  Run in the build container:

    python tools/make_golden_unfrozen.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import reference_harness as rh  # noqa: E402
import interactron_b200 as ib  # noqa: E402
from interactron_b200.synthetic import synthetic_episode  # noqa: E402


def main():
    cfg = ib.default_config("interactron_random", weights="synthetic")
    mine = ib.build_model(cfg.MODEL)
    ref = rh.build_reference_model("interactron_random", mine.state_dict(), freeze_backbone=False)
    names = rh.reference_fast_weight_names(ref)
    ep = 0
    tr = rh.reference_predict_with_trace(ref, synthetic_episode(ep))
    gold = {"episode": ep, "n_theta": len(names), "n_theta_backbone": sum("backbone" in n for n in names),
            "pred_logits": tr["out"]["pred_logits"], "pred_boxes": tr["out"]["pred_boxes"],
            "learned_loss": tr["learned_loss"].clone()}
    # the same call in D1 mode, for the size of the gap
    ref_d1 = rh.build_reference_model("interactron_random", mine.state_dict(), freeze_backbone=True)
    d1 = rh.reference_predict_with_trace(ref_d1, synthetic_episode(ep))
    rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
    gold["d1_gap_logits"] = rel(d1["out"]["pred_logits"], gold["pred_logits"])
    gold["d1_gap_boxes"] = rel(d1["out"]["pred_boxes"], gold["pred_boxes"])
    print(f"theta: {gold['n_theta']} tensors ({gold['n_theta_backbone']} backbone); D1 vs unmodified reference: "
          f"logits {gold['d1_gap_logits']:.3e}, boxes {gold['d1_gap_boxes']:.3e}")
    torch.save(gold, os.path.join(ROOT, "tests", "golden", "interactron_random_predict_unfrozen.pt"))


if __name__ == "__main__":
    main()
