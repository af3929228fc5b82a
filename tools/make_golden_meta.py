"""Generates tests/golden/<model>_forward.pt: the UNMODIFIED reference's meta-training step
`forward(data)` (reference models/interactron.py:61-151, interactron_random.py:57-136; D1 mode,
eval()) on seeded synthetic weights / episodes, run on CPU in fp32 (the reference as it ships) and in
fp64 (how far fp32 itself is from the exact gradients).  Per parameter the fixture keeps the gradient
norm and a strided sample of 256..511 elements (stride = (numel // 256) | 1; everything for small tensors).

    python tools/make_golden_meta.py            (build container only: needs /root/reference)
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import reference_harness as rh  # noqa: E402
import interactron_b200 as ib  # noqa: E402
from interactron_b200.synthetic import collate_episodes, synthetic_episode  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
EPISODES, RIDX = (0, 1), (3, 1)


def stride_of(numel):
    return max(1, numel // 256) | 1          # odd: walks through the columns of power-of-two matrices


def pack(grads):
    out = {}
    for n, g in grads.items():
        if g is None:
            out[n] = None
            continue
        f = g.reshape(-1)
        out[n] = {"norm": f.double().norm(), "sample": f[::stride_of(f.numel())].clone()}
    return out


def run(model_type):
    cfg = ib.default_config(model_type, weights="synthetic")
    mine = ib.build_model(cfg.MODEL)
    data = collate_episodes([synthetic_episode(e) for e in EPISODES])
    gold = {"episodes": EPISODES, "ridx": RIDX}
    for tag, dt in (("fp32", torch.float32), ("fp64", torch.float64)):
        ref = rh.build_reference_model(model_type, mine.state_dict()).to(dt)
        d = dict(data)
        d["frames"] = data["frames"].to(dt)
        d["boxes"] = [[b.to(dt) for b in ep] for ep in data["boxes"]]
        torch.set_default_dtype(dt)
        try:
            rounds = []
            p, l, g = rh.reference_forward_with_grads(ref, d, RIDX)
            rounds.append({"predictions": p, "losses": l, "grads": pack(g)})
        finally:
            torch.set_default_dtype(torch.float32)
        gold[tag] = rounds
        print(model_type, tag, {k: float(v) for k, v in rounds[-1]["losses"].items()}, flush=True)
    torch.save(gold, os.path.join(OUT, f"{model_type}_forward.pt"))


if __name__ == "__main__":
    for m in sys.argv[1:] or ("interactron_random", "interactron"):
        run(m)
