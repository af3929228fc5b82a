"""Golden for padded frames: the UNMODIFIED reference's predict() (D1 mode) on an episode whose frames
1 and 3 are padded on the right / bottom (mask != 0, pixels zeroed) - the case the dataset never
produces (datasets/sequence_dataset.py:56 writes all-zero masks) but the interface allows: the mask
reaches the key-padding mask of every encoder / cross attention and the sine position embedding
(detr_models/backbone.py:77, position_encoding.py:28-48).  -> tests/golden/{interactron_random,interactron}_predict_masked.pt

    python tools/make_golden_masked.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import reference_harness as rh  # noqa: E402
import interactron_b200 as ib  # noqa: E402
from interactron_b200.synthetic import masked_episode  # noqa: E402

def make(model_type, episodes):
    cfg = ib.default_config(model_type, weights="synthetic")
    mine = ib.build_model(cfg.MODEL)
    ref = rh.build_reference_model(model_type, mine.state_dict())
    gold = {}
    for ep in episodes:
        data = masked_episode(ep)
        tr = rh.reference_predict_with_trace(ref, data)
        gold[ep] = {"pred_logits": tr["out"]["pred_logits"].clone(), "pred_boxes": tr["out"]["pred_boxes"].clone(),
                    "pre_logits": tr["pre"]["pred_logits"][0].clone(), "pre_boxes": tr["pre"]["pred_boxes"][0].clone(),
                    "learned_loss": tr["learned_loss"].clone(),
                    "g_norms": torch.stack([g.norm() for g in tr["grads"]])}
        print(model_type, "episode", ep, "learned_loss", float(tr["learned_loss"]), "masked px", int(data["masks"].sum()))
    torch.save(gold, os.path.join(ROOT, "tests", "golden", f"{model_type}_predict_masked.pt"))


if __name__ == "__main__":
    make("interactron_random", (3,))
    make("interactron", (3,))
