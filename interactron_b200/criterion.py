"""HungarianMatcher and SetCriterion with the reference's interface (reference
models/detr_models/matcher.py:12-77, models/detr_models/detr.py:86-265), computing on the device:

  matcher   cost matrix in `itn_matcher_cost` (block diagonal only - the reference builds the full
            [F*Q, sum T] matrix and uses its diagonal blocks), one D2H copy, LSAP with the same
            scipy.optimize.linear_sum_assignment on the same fp32 costs -> identical indices.
  criterion `itn_criterion`: weighted CE, L1, GIoU, class_error, cardinality_error in five small
            launches with fixed-order reductions; `loss_and_grad` also returns the gradient of the
            reference's scalarisation (models/interactron.py:121-122) wrt logits and boxes.

Both accept the reference's `outputs` / `targets` structures.  CUDA only (no CPU path).
"""
import torch
from scipy.optimize import linear_sum_assignment
from torch import nn

LOSS_KEYS = ("loss_ce", "class_error", "cardinality_error", "loss_bbox", "loss_giou")


def _ops_for(t):
    from .ops import CudaOps
    if not t.is_cuda:
        raise RuntimeError("interactron_b200.criterion runs on CUDA tensors only (there is no CPU path)")
    key = t.device.index
    ops = _OPS.get(key)
    if ops is None:
        ops = _OPS[key] = CudaOps(t.device)
    return ops


_OPS = {}


def _pack_targets(targets, device, dtype=torch.float32, num_logits=None):
    """list of {"labels": [T_f], "boxes": [T_f,4]} -> concatenated labels/boxes + int32 offsets.
    num_logits: width of the logits rows the labels index (matcher cost / criterion kernels read
    logits[label]): a label outside [0, num_logits) raises IndexError here, as torch's gather / cross_entropy
    does in the reference (models/detr_models/matcher.py:61, detr.py:126), instead of an out-of-bounds read."""
    sizes = [int(t["labels"].numel()) for t in targets]
    off = [0]
    for n in sizes:
        off.append(off[-1] + n)
    if off[-1]:
        labels = torch.cat([t["labels"].reshape(-1) for t in targets]).to(device=device, dtype=torch.int64)
        boxes = torch.cat([t["boxes"].reshape(-1, 4) for t in targets]).to(device=device, dtype=dtype)
        if num_logits is not None:
            lo, hi = int(labels.min()), int(labels.max())          # one min/max over all labels of the call
            if lo < 0 or hi >= num_logits:
                raise IndexError(f"target label out of range: labels span [{lo}, {hi}] but the logits have "
                                 f"{num_logits} classes")
    else:
        labels = torch.zeros(0, dtype=torch.int64, device=device)
        boxes = torch.zeros(0, 4, dtype=dtype, device=device)
    return labels.contiguous(), boxes.contiguous(), torch.tensor(off, dtype=torch.int32, device=device), sizes, off


class HungarianMatcher(nn.Module):
    def __init__(self, cost_class=1.0, cost_bbox=1.0, cost_giou=1.0):
        super().__init__()
        assert cost_class != 0 or cost_bbox != 0 or cost_giou != 0, "all costs cant be 0"
        self.cost_class, self.cost_bbox, self.cost_giou = cost_class, cost_bbox, cost_giou
        self._ops_override = None       # tests only: the torch simulation of the kernel interface

    @torch.no_grad()
    def forward(self, outputs, targets):
        """-> [(index_i int64, index_j int64)] per frame, as matcher.py:75-77."""
        dt = torch.float32 if self._ops_override is None else outputs["pred_logits"].dtype
        logits = outputs["pred_logits"].detach().to(dt).contiguous()
        boxes = outputs["pred_boxes"].detach().to(dt).contiguous()
        Fn, Q = logits.shape[:2]
        ops = self._ops_override or _ops_for(logits)
        labels, tboxes, off_t, sizes, off = _pack_targets(targets, logits.device, logits.dtype, logits.shape[-1])
        if off[-1] == 0:
            e = torch.zeros(0, dtype=torch.int64)
            return [(e, e) for _ in range(Fn)]
        cost = ops.matcher_cost(logits, boxes, tboxes, labels, off_t, self.cost_class, self.cost_bbox,
                                self.cost_giou).cpu()
        out = []
        for f in range(Fn):
            c = cost[Q * off[f]:Q * off[f + 1]].view(Q, sizes[f])
            i, j = linear_sum_assignment(c.numpy())
            out.append((torch.as_tensor(i, dtype=torch.int64), torch.as_tensor(j, dtype=torch.int64)))
        return out


class SetCriterion(nn.Module):
    """Same constructor and `forward(outputs, targets, detector_out=None, background_c=0.1)` as the
    reference; `empty_weight` is registered for state_dict compatibility only (the reference ignores
    it too: loss_labels rebuilds the weight from background_c, detr.py:124-126)."""

    def __init__(self, num_classes, matcher, weight_dict=None, eos_coef=0.1, losses=("labels", "boxes", "cardinality")):
        super().__init__()
        self.num_classes = num_classes
        self.matcher = matcher
        self.weight_dict = weight_dict or {"loss_ce": 1, "loss_bbox": 5, "loss_giou": 2}
        self.eos_coef = eos_coef
        self.losses = list(losses)
        w = torch.ones(num_classes + 1)
        w[-1] = eos_coef
        self.register_buffer("empty_weight", w)
        self._ops_override = None       # tests only: the torch simulation of the kernel interface

    def _run(self, outputs, targets, indices, background_c, groups, weights, want_grad):
        logits = outputs["pred_logits"]
        boxes = outputs["pred_boxes"]
        lead = logits.shape[:-2]
        dt = torch.float32 if self._ops_override is None else logits.dtype
        logits = logits.detach().to(dt).reshape(-1, *logits.shape[-2:]).contiguous()
        boxes = boxes.detach().to(dt).reshape(-1, *boxes.shape[-2:]).contiguous()
        Fn, Q, Cn = logits.shape
        if Cn != self.num_classes + 1:
            raise ValueError(f"pred_logits has {Cn} classes, criterion built for {self.num_classes}+1")
        if len(targets) != Fn or len(indices) != Fn or Fn % groups:
            raise ValueError("targets / indices must have one entry per frame")
        ops = self._ops_override or _ops_for(logits)
        labels, tboxes, off_t, sizes, off = _pack_targets(targets, logits.device, logits.dtype, logits.shape[-1])
        rows, tg, moff = [], [], [0]
        per_group = Fn // groups
        for f, (i, j) in enumerate(indices):
            rows.append(i.to(torch.int64) + f * Q)
            tg.append(j.to(torch.int64) + off[f])
            if (f + 1) % per_group == 0:
                moff.append(sum(int(r.numel()) for r in rows))
        dev = logits.device
        match_row = torch.cat(rows).to(device=dev, dtype=torch.int32)
        match_tgt = torch.cat(tg).to(device=dev, dtype=torch.int32)
        match_off = torch.tensor(moff, dtype=torch.int32, device=dev)
        res = ops.criterion(logits, boxes, tboxes, labels, off_t, match_row, match_tgt, match_off, groups,
                            background_c, weights, want_grad)
        if want_grad:
            losses, dl, db = res
            return losses, dl.reshape(*lead, Q, Cn), db.reshape(*lead, Q, 4)
        return res

    def _indices(self, outputs, targets, detector_out):
        src = detector_out if detector_out is not None else outputs
        flat = {k: src[k].reshape(-1, *src[k].shape[-2:]) for k in ("pred_logits", "pred_boxes")}
        return self.matcher(flat, targets)

    def forward(self, outputs, targets, detector_out=None, background_c=0.1):
        """-> {"loss_ce", "class_error", "loss_bbox", "loss_giou", "cardinality_error"}: 0-dim tensors,
        in the reference's key order (labels, boxes, cardinality; detr.py:247-249)."""
        indices = self._indices(outputs, targets, detector_out)
        v = self._run(outputs, targets, indices, background_c, 1, (1.0, 1.0, 1.0), False)[0]
        d = dict(zip(LOSS_KEYS, v.unbind(0)))
        order = {"labels": ("loss_ce", "class_error"), "boxes": ("loss_bbox", "loss_giou"),
                 "cardinality": ("cardinality_error",)}
        return {k: d[k] for name in self.losses for k in order[name]}

    def group_losses(self, outputs, targets, background_c=0.1, groups=1, detector_out=None):
        """`groups` independent criterion calls in one launch sequence -> [groups, 5] (LOSS_KEYS order)."""
        indices = self._indices(outputs, targets, detector_out)
        return self._run(outputs, targets, indices, background_c, groups, (1.0, 1.0, 1.0), False)

    def loss_and_grad(self, outputs, targets, background_c=0.1, groups=1, weights=(1.0, 2.0, 5.0),
                      detector_out=None):
        """Batched over `groups` criterion calls.  -> (losses [groups,5] in LOSS_KEYS order, dlogits,
        dboxes) where the gradients are those of  w[0]*loss_ce + w[1]*loss_bbox + w[2]*loss_giou  summed
        over the groups.  The default weights are the scalarisation the adaptive models back-propagate,
        `ce + 5*giou + 2*bbox` (models/interactron.py:121-122) - NOT weight_dict."""
        indices = self._indices(outputs, targets, detector_out)
        return self._run(outputs, targets, indices, background_c, groups, weights, True)


def build_criterion(cfg):
    """reference models/detr_models/detr.py:334-338."""
    matcher = HungarianMatcher(cost_class=cfg.SET_COST_CLASS, cost_bbox=cfg.SET_COST_BBOX, cost_giou=cfg.SET_COST_GIOU)
    return SetCriterion(cfg.NUM_CLASSES, matcher=matcher, weight_dict={"loss_ce": 1, "loss_bbox": 5, "loss_giou": 2},
                        eos_coef=0.1, losses=["labels", "boxes", "cardinality"])
