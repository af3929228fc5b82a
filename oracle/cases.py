"""Seeded synthetic inputs shared by tools/make_golden.py and tests/ (test infrastructure, like the rest
of oracle/: never imported by the product path)."""
import torch


def criterion_case(seed, frames=5, queries=50, classes=1236):
    """Predictions + targets for SetCriterion / HungarianMatcher: ~70 % of the queries favour
    "no object", one query per target is pulled towards it (label logit +6, box = target + noise), so that
    class_error, cardinality_error and the box losses are all non-trivial."""
    gen = torch.Generator().manual_seed(7000 + 31 * seed + frames)
    logits = torch.randn(frames, queries, classes, generator=gen)
    boxes = torch.cat([torch.rand(frames, queries, 2, generator=gen) * 0.6 + 0.2,
                       torch.rand(frames, queries, 2, generator=gen) * 0.3 + 0.05], -1)
    noobj = torch.rand(frames, queries, generator=gen) < 0.7
    logits[..., -1] += 6.0 * noobj
    targets = []
    for f in range(frames):
        n = int(torch.randint(0 if seed == 2 and f == 1 else 3, 9, (1,), generator=gen))
        labels = torch.randint(1, classes - 1, (n,), generator=gen)
        tb = torch.cat([torch.rand(n, 2, generator=gen) * 0.6 + 0.2, torch.rand(n, 2, generator=gen) * 0.3 + 0.05], 1)
        qs = torch.randperm(queries, generator=gen)[:n]
        for t in range(n):
            if t % 3 != 2:                       # every third target stays unclaimed by design
                logits[f, qs[t], labels[t]] += 6.0
                boxes[f, qs[t]] = (tb[t] + 0.03 * torch.randn(4, generator=gen)).clamp(0.02, 0.98)
        targets.append({"labels": labels, "boxes": tb})
    return logits, boxes, targets
