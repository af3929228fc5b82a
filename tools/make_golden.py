"""Generates tests/golden/*.pt by running the UNMODIFIED reference (/root/reference, CPU, D1 mode:
backbone frozen) on the seeded synthetic weights and episodes.  Run in the build container:

    python tools/make_golden.py

The fixtures let the GPU parity tests run on a box where /root/reference does not exist.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import reference_harness as rh  # noqa: E402
import interactron_b200 as ib  # noqa: E402
from interactron_b200.synthetic import synthetic_episode  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
SMALL = ("norm", "bias", "query_embed")          # theta tensors stored in full (small ones)


def adaptive(model_type, episodes):
    cfg = ib.default_config(model_type, weights="synthetic")
    mine = ib.build_model(cfg.MODEL)
    ref = rh.build_reference_model(model_type, mine.state_dict())
    names = rh.reference_fast_weight_names(ref)
    gold = {"theta_names": names, "episodes": {}}
    for ep in episodes:
        data = synthetic_episode(ep)
        tr = rh.reference_predict_with_trace(ref, data)
        small = {n: (g.clone(), tp.clone()) for n, g, tp in zip(names, tr["grads"], tr["theta_prime"])
                 if any(s in n for s in SMALL) and g.numel() <= 12800}
        gold["episodes"][ep] = {
            "pred_logits": tr["out"]["pred_logits"], "pred_boxes": tr["out"]["pred_boxes"],
            "box_features": tr["out"]["box_features"],
            "pre_logits_f0": tr["pre"]["pred_logits"][0, 0].clone(),
            "pre_logits_f4": tr["pre"]["pred_logits"][0, 4].clone(),
            "pre_boxes": tr["pre"]["pred_boxes"][0].clone(),
            "loss_vec": tr["fusion_loss"].reshape(-1).clone(),
            "learned_loss": tr["learned_loss"].clone(),
            "actions": tr["fusion_actions"].clone(),
            "g_norms": torch.stack([g.norm() for g in tr["grads"]]),
            "theta_prime_norms": torch.stack([p.double().norm() for p in tr["theta_prime"]]),
            "theta_step_norms": torch.stack([(p - q).norm() for p, q in
                                             zip(tr["theta_prime"], [dict(ref.detector.named_parameters())[n]
                                                                     for n in names])]),
            "small": small,
        }
        print(model_type, "episode", ep, "learned_loss", float(tr["learned_loss"]))
    torch.save(gold, os.path.join(OUT, f"{model_type}_predict.pt"))
    return mine, ref


def baselines():
    gold = {}
    cfg = ib.default_config("single_frame_baseline", weights="synthetic")
    mine = ib.build_model(cfg.MODEL)
    ref = rh.build_reference_model("detr", {("detector." + k[len("model."):] if k.startswith("model.") else k): v
                                            for k, v in mine.state_dict().items()})
    data = synthetic_episode(0, frames=1)
    with torch.no_grad():
        o = ref.predict(data)
    gold["detr_ep0_1frame"] = {k: o[k].clone() for k in ("pred_logits", "pred_boxes", "box_features")}
    cfg = ib.default_config("multi_frame_baseline", weights="synthetic")
    mine = ib.build_model(cfg.MODEL)
    ref = rh.build_reference_model("detr_multiframe", mine.state_dict())
    data = synthetic_episode(0)
    with torch.no_grad():
        o = ref.predict(data)
    gold["detr_multiframe_ep0"] = {k: o[k].clone() for k in ("pred_logits", "pred_boxes")}
    torch.save(gold, os.path.join(OUT, "baselines_predict.pt"))
    print("baselines done")


def policy():
    """get_next_action on 1..4 frames (reference models/interactron.py:174-197)."""
    cfg = ib.default_config("interactron", weights="synthetic")
    mine = ib.build_model(cfg.MODEL)
    ref = rh.build_reference_model("interactron", mine.state_dict())
    gold = {}
    data = synthetic_episode(0)
    for s in range(1, 5):
        d = dict(data)
        d["frames"], d["masks"] = data["frames"][:, :s], data["masks"][:, :s]
        d["category_ids"] = [data["category_ids"][0][:s]]
        d["boxes"] = [data["boxes"][0][:s]]
        with torch.no_grad():
            gold[s] = ref.get_next_action(d)
    torch.save(gold, os.path.join(OUT, "interactron_actions.pt"))
    print("policy actions", gold)


def matcher():
    """Reference HungarianMatcher on seeded synthetic predictions/targets."""
    rh._load()
    from models.detr_models.matcher import HungarianMatcher
    m = HungarianMatcher(cost_class=1, cost_bbox=5, cost_giou=2)
    gold = {}
    for seed in (0, 1, 2):
        gen = torch.Generator().manual_seed(100 + seed)
        logits = torch.randn(5, 50, 1236, generator=gen)
        boxes = torch.rand(5, 50, 4, generator=gen) * 0.5 + 0.1
        tg = []
        for f in range(5):
            n = int(torch.randint(3, 9, (1,), generator=gen))
            tg.append({"labels": torch.randint(1, 1235, (n,), generator=gen),
                       "boxes": torch.cat([torch.rand(n, 2, generator=gen) * 0.6 + 0.2,
                                           torch.rand(n, 2, generator=gen) * 0.3 + 0.05], 1)})
        idx = m({"pred_logits": logits, "pred_boxes": boxes}, tg)
        gold[seed] = [(i.clone(), j.clone()) for i, j in idx]
    torch.save(gold, os.path.join(OUT, "matcher_assignments.pt"))
    print("matcher done")


def criterion():
    """Reference SetCriterion (+ autograd of ce + 5*giou + 2*bbox, models/interactron.py:121-122) on
    oracle/cases.criterion_case inputs."""
    rh._load()
    from models.detr_models.detr import SetCriterion
    from models.detr_models.matcher import HungarianMatcher
    from oracle.cases import criterion_case
    crit = SetCriterion(1235, matcher=HungarianMatcher(1, 5, 2), weight_dict={"loss_ce": 1, "loss_bbox": 5, "loss_giou": 2},
                        eos_coef=0.1, losses=["labels", "boxes", "cardinality"])
    gold = {}
    for seed in (0, 1, 2):
        for frames in (5, 1):
            logits, boxes, targets = criterion_case(seed, frames)
            logits.requires_grad_(True)
            boxes.requires_grad_(True)
            out = crit({"pred_logits": logits, "pred_boxes": boxes}, targets, background_c=0.1)
            total = out["loss_ce"] + 5 * out["loss_giou"] + 2 * out["loss_bbox"]
            dl, db = torch.autograd.grad(total, (logits, boxes))
            idx = crit.matcher({"pred_logits": logits.detach(), "pred_boxes": boxes.detach()}, targets)
            gold[(seed, frames)] = {
                "losses": {k: v.detach().clone() for k, v in out.items()}, "keys": list(out.keys()),
                "indices": [(i.clone(), j.clone()) for i, j in idx], "dboxes": db.clone(),
                "dlogits_cols": dl[..., ::97].clone(), "dlogits_abs_rowsum": dl.abs().sum(-1).clone()}
    torch.save(gold, os.path.join(OUT, "criterion.pt"))
    print("criterion done", {k: {n: round(float(x), 5) for n, x in v["losses"].items()} for k, v in gold.items()})


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    adaptive("interactron_random", (0, 1))
    adaptive("interactron", (0,))
    baselines()
    policy()
    matcher()
    criterion()
