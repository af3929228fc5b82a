"""Golden vectors of the policy step `get_next_action` (reference models/interactron.py:174-197): the ACTION
LOGITS the fusion network emits (not only their argmax) for several synthetic episodes with 1..4 frames seen,
from the unmodified reference on CPU.  -> tests/golden/interactron_action_logits.pt

    python tools/make_golden_policy.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import reference_harness as rh  # noqa: E402
import interactron_b200 as ib  # noqa: E402
from interactron_b200.synthetic import synthetic_episode  # noqa: E402

EPISODES = (0, 3, 7, 12)

cfg = ib.default_config("interactron", weights="synthetic")
mine = ib.build_model(cfg.MODEL)
ref = rh.build_reference_model("interactron", mine.state_dict())
stash = []
fusion_forward = ref.fusion.forward


def tap(x):
    out = fusion_forward(x)
    stash.append(out["actions"].detach().clone())
    return out


ref.fusion.forward = tap
gold = {"episodes": EPISODES, "actions": {}, "logits": {}}
for ep in EPISODES:
    data = synthetic_episode(ep)
    for s in range(1, 5):
        d = dict(data)
        d["frames"], d["masks"] = data["frames"][:, :s], data["masks"][:, :s]
        d["category_ids"] = [data["category_ids"][0][:s]]
        d["boxes"] = [data["boxes"][0][:s]]
        with torch.no_grad():
            a = ref.get_next_action(d)
        lg = stash.pop()
        assert int(lg[s - 1].argmax(-1)) == a
        gold["actions"][(ep, s)] = a
        gold["logits"][(ep, s)] = lg.reshape(-1, lg.shape[-1])[: max(s, 4)].clone()
        print(ep, s, a, [round(float(v), 5) for v in lg[s - 1].flatten()])
torch.save(gold, os.path.join(ROOT, "tests", "golden", "interactron_action_logits.pt"))
