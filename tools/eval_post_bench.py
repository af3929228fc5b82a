"""Evaluator post-processing throughput: images/s of the reference's per-image torch sequence
run on the device (softmax-max, filter, nms, per-category matching with its .item() round trips =
oracle/port.evaluator_records on CUDA tensors) vs DetectionPostprocessor + image_detections."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interactron_b200.evaluator import DetectionPostprocessor, image_detections  # noqa: E402
from interactron_b200.ops import CudaOps  # noqa: E402
from oracle import cases, port  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
ops = CudaOps()
batch = [cases.evaluator_case(s) for s in range(B)]
logits = torch.stack([b[0] for b in batch]).cuda()
boxes = torch.stack([b[1] for b in batch]).cuda()
ids = list(range(1, 1235))


class Holder:
    def _get_ops(self):
        return ops


pp = DetectionPostprocessor(Holder())
preds = {"pred_logits": logits[:, None], "pred_boxes": boxes[:, None]}


def ours():
    post = pp(preds)
    return [image_detections(post, i, batch[i][2], batch[i][3], "x", ids) for i in range(B)]


def theirs():
    out = []
    for i in range(B):
        out.append(port.evaluator_records(logits[i], boxes[i], batch[i][2].cuda(), batch[i][3].cuda(), "x", ids))
    return out


def timed(fn, n=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n


e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ops.detect_postprocess(logits, boxes, 1235, 0.5)
e0.record()
for _ in range(20):
    ops.detect_postprocess(logits, boxes, 1235, 0.5)
e1.record()
torch.cuda.synchronize()
kern_ms = e0.elapsed_time(e1) / 20
t_ours, t_ref = timed(ours), timed(theirs)
print(json.dumps({"images": B, "kernel_ms": kern_ms, "kernel_GBs": logits.numel() * 4 / kern_ms / 1e6,
                  "ours_images_per_s": B / t_ours, "reference_sequence_on_device_images_per_s": B / t_ref}))
