"""Evaluator post-processing of predict()'s output.

The reference evaluators (engine/random_policy_evaluator.py:61-158, engine/interactive_evaluator.py:67-164)
turn every image's `pred_logits / pred_boxes` into TP / FP / FN records with ~10 device round trips and
dozens of `.item()` synchronisations per image.  Here the device part - softmax-max, background filter,
cxcywh->xyxy, class-agnostic NMS - is ONE kernel launch for all images of a predict() batch
(`itn_detect_postprocess`, csrc/itn_postprocess.cu), its five small outputs are read back once, and the
record building runs on host arrays with no further synchronisation.  Records are the reference's:
same keys, same order (its Python-set iteration order included), same float32 values.
"""
import numpy as np
import torch

BACKGROUND = 1235          # reference: `pred_cats != 1235` (random_policy_evaluator.py:70)


class DetectionPostprocessor:
    def __init__(self, model, background=BACKGROUND, iou_threshold=0.5):
        self.model, self.background, self.iou_threshold = model, int(background), float(iou_threshold)

    def __call__(self, predictions):
        """predictions: predict()'s dict (`pred_logits [B,1,Q,C]`, `pred_boxes [B,1,Q,4]`, on the device)
        -> host arrays {count [B], keep_idx [B,Q], score [B,Q], cat [B,Q], xyxy [B,Q,4]}; entries
        k < count[b] are image b's detections after NMS, in decreasing score order."""
        ops = self.model._get_ops()
        logits = predictions["pred_logits"][:, 0].contiguous()
        boxes = predictions["pred_boxes"][:, 0].contiguous()
        out = ops.detect_postprocess(logits, boxes, self.background, self.iou_threshold)
        packed = [t.cpu() for t in out]                     # the only device->host traffic (a few KB per image)
        names = ("count", "keep_idx", "score", "cat", "xyxy")
        return {n: t.numpy() for n, t in zip(names, packed)}


def cxcywh_to_xyxy(b):
    """detr_models/util/box_ops.py:8-12 on a float32 array [n,4]."""
    b = np.asarray(b, dtype=np.float32).reshape(-1, 4)
    half = np.float32(0.5)
    return np.stack([b[:, 0] - half * b[:, 2], b[:, 1] - half * b[:, 3],
                     b[:, 0] + half * b[:, 2], b[:, 1] + half * b[:, 3]], axis=1)


def box_iou(a, b):
    """torchvision.ops.box_iou in float32, same operation order."""
    a, b = np.asarray(a, np.float32).reshape(-1, 4), np.asarray(b, np.float32).reshape(-1, 4)
    area_a = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    area_b = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    lt = np.maximum(a[:, None, :2], b[None, :, :2])
    rb = np.minimum(a[:, None, 2:], b[None, :, 2:])
    wh = np.clip(rb - lt, 0, None)
    inter = wh[..., 0] * wh[..., 1]
    return inter / (area_a[:, None] + area_b[None, :] - inter)


def match_predictions(ious):
    """The reference's proposal rounds between predictions (rows) and ground truths (columns)
    (utils/detection_utils.py:401-421): every still-free prediction proposes to its next-best ground
    truth; a ground truth keeps, among this round's proposers, the one with the largest IoU (dropping
    whoever it held); stops once min(#pred, #gt) predictions are held.  -> (best_iou [g], best_pred [g]),
    -1 / 0.0 where the ground truth ends with no overlapping prediction.
    torch.argsort / argmax on CPU tensors keep the reference's tie-breaking."""
    t = torch.from_numpy(np.ascontiguousarray(ious, dtype=np.float32))
    p, g = t.shape
    pref = torch.argsort(t, dim=1, descending=True)
    nxt = torch.zeros(p, dtype=torch.long)
    free = torch.ones(p, dtype=torch.bool)
    held = -torch.ones(g, dtype=torch.long)
    rows = torch.arange(p)
    for _ in range(g):
        proposal = pref[rows, nxt]
        for j in range(g):
            winner = torch.argmax(t[:, j] * (proposal == j))
            if held[j] != -1 and held[j] != winner:
                free[held[j]] = True
            held[j] = winner
            free[winner] = False
        nxt[free] += 1
        if int(torch.count_nonzero(~free)) >= min(p, g):
            break
    best = torch.zeros(g)
    has = held != -1
    best[has] = t[held[has], has]
    held[best == 0.0] = -1
    return best.numpy(), held.numpy()


def image_detections(post, b, gt_boxes, gt_cats, img, class_ids=None):
    """The reference's per-image record list (random_policy_evaluator.py:79-146) from the host arrays of
    DetectionPostprocessor.  gt_boxes [n,4] cxcywh, gt_cats [n]; class_ids = the reference's
    THOR_CLASS_IDS (utils/constants.py:173; default: every foreground id)."""
    n = int(post["count"][b])
    pred_boxes = post["xyxy"][b, :n]
    pred_scores = post["score"][b, :n]
    pred_cats = post["cat"][b, :n].astype(np.int64)
    gt_boxes = cxcywh_to_xyxy(gt_boxes.cpu().numpy() if torch.is_tensor(gt_boxes) else gt_boxes)
    gt_cats = (gt_cats.cpu().numpy() if torch.is_tensor(gt_cats) else np.asarray(gt_cats)).astype(np.int64)
    class_ids = range(1, BACKGROUND) if class_ids is None else class_ids

    def record(kind, match, cat, iou, score, box):
        area = (box[2] - box[0]) * (box[3] - box[1])               # float32 arithmetic, as in the reference
        return {"iou": float(iou), "category_match": match, "type": kind, "pred_cat": cat, "pred_score": float(score),
                "box": [float(c) for c in box], "area": float(area), "img": img}

    pred_cat_set = set([int(c) for c in pred_cats])
    gt_cat_set = set([int(c) for c in gt_cats])
    pred_only = set(class_ids).intersection(pred_cat_set - gt_cat_set)
    dets = []
    for cat in gt_cat_set:
        mine, theirs = pred_cats == cat, gt_cats == cat
        if mine.any():
            pb, ps, gb = pred_boxes[mine], pred_scores[mine], gt_boxes[theirs]
            ious = box_iou(pb, gb)
            best_iou, best_pred = match_predictions(ious)
            for i in range(ious.shape[0]):
                kind = "tp" if (best_pred == i).any() else "fp"
                dets.append(record(kind, True, cat, ious[i].max(), ps[i], pb[i]))
            for j in range(ious.shape[1]):
                if best_iou[j] == 0.0:
                    dets.append(record("fn", False, cat, 0.0, 0.0, gb[j]))
        else:
            for box in gt_boxes[theirs]:
                dets.append(record("fn", False, cat, 0.0, 0.0, box))
    for cat in pred_only:
        mine = pred_cats == cat
        for box, sc in zip(pred_boxes[mine], pred_scores[mine]):
            dets.append(record("fp", False, cat, 0.0, sc, box))
    return dets
