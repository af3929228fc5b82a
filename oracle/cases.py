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


def evaluator_case(seed, queries=50, classes=1236):
    """One image's predictions + ground truth for the evaluator post-processing (TP / FP / FN records):
    3-6 ground-truth boxes from a pool of 3 categories (so categories repeat); every ground truth is claimed
    by 0-3 queries (label logit +6..+10, box = ground truth + jitter, so NMS has duplicates to remove);
    a few queries fire on a category that is not in the image; the rest prefer background (class
    `classes - 1`).  seed % 4 == 3: no ground truth at all."""
    gen = torch.Generator().manual_seed(9100 + seed)
    logits = torch.randn(queries, classes, generator=gen)
    boxes = torch.cat([torch.rand(queries, 2, generator=gen) * 0.6 + 0.2, torch.rand(queries, 2, generator=gen) * 0.3 + 0.05], -1)
    logits[:, -1] += 9.0
    pool = torch.randint(1, classes - 1, (3,), generator=gen)
    n = 0 if seed % 4 == 3 else int(torch.randint(3, 7, (1,), generator=gen))
    gt_cats = pool[torch.randint(0, 3, (n,), generator=gen)]
    gt_boxes = torch.cat([torch.rand(n, 2, generator=gen) * 0.6 + 0.2, torch.rand(n, 2, generator=gen) * 0.3 + 0.05], 1)
    order = torch.randperm(queries, generator=gen).tolist()
    for t in range(n):
        for _ in range(int(torch.randint(0, 4, (1,), generator=gen))):
            q = order.pop()
            logits[q, -1] -= 9.0
            logits[q, gt_cats[t]] += 6.0 + 4.0 * float(torch.rand(1, generator=gen))
            boxes[q] = (gt_boxes[t] + 0.02 * torch.randn(4, generator=gen)).clamp(0.02, 0.98)
    for _ in range(3):                                   # detections of categories absent from the image
        q = order.pop()
        logits[q, -1] -= 9.0
        logits[q, int(torch.randint(1, classes - 1, (1,), generator=gen))] += 8.0
    return logits, boxes, gt_boxes, gt_cats
