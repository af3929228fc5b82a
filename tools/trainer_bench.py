"""Roofline of the trainer step: itn_sumsq_partials + 3x itn_clip_adam_step over the flat
buffers of `interactron` (58.4 M elements) vs torch's clip_grad_norm_ + 2x Adam on the same device.
Algorithmic bytes: 4 B/elem (norm pass) + 32 B/elem (read g,w,m,v; write w,m,v,g=0)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import interactron_b200 as ib  # noqa: E402
from interactron_b200 import meta  # noqa: E402
from interactron_b200.trainer import MetaTrainerStep  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "interactron"
model = ib.build_model(ib.default_config(name, weights="synthetic").MODEL).cuda().eval()
import copy  # noqa: E402
ref = copy.deepcopy(model)
tr = MetaTrainerStep(model, 1e-5, 1e-4, 1.0)
loop = model._get_loop()
sizes = (loop.theta_pack.numel, loop.psi_pack.numel, loop.phi_pack.numel)
n = sum(sizes)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def install(G):
    flat = {"all": G, "theta": G[:, :sizes[0]], "psi": G[:, sizes[0]:sizes[0] + sizes[1]], "phi": G[:, sizes[0] + sizes[1]:]}
    model.last_meta_grads = flat
    meta.accumulate_grads(model, flat)


def timed(fn, prep, iters=10):
    ms = []
    for i in range(iters + 3):
        prep()
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ms.append(e0.elapsed_time(e1))
    ms.sort()
    return ms[len(ms) // 2]


ours = timed(lambda: tr.step(), lambda: install(torch.randn(1, n, device="cuda") * 1e-3))
Gk = torch.randn(n, device="cuda") * 1e-3
kern = timed(lambda: tr._kernels(Gk), lambda: None)
l0 = loop.ops.launch_count()
install(torch.randn(1, n, device="cuda") * 1e-3)
tr.step()
launches = loop.ops.launch_count() - l0
opt_d = torch.optim.Adam(ref.detector.parameters(), lr=1e-5)
opt_s = torch.optim.Adam(ref.fusion.parameters(), lr=1e-4)
with_grad = [p for nm, p in ref.named_parameters() if not nm.startswith("detector.backbone")]


def ref_prep():
    for p in with_grad:
        p.grad = torch.randn_like(p) * 1e-3


def ref_step():
    torch.nn.utils.clip_grad_norm_(ref.parameters(), 1.0)
    opt_d.step(); opt_s.step(); opt_d.zero_grad(); opt_s.zero_grad()


theirs = timed(ref_step, ref_prep)
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
gb = 36.0 * n / 1e9
print(json.dumps({"workload": f"{name} trainer step (clip_grad_norm_ + 2x Adam + zero_grad)", "elements": n,
                  "kernels_ms": kern, "step_ms": ours, "itn_launches": launches, "includes": "W^T twin rebuild of every weight (transposes)",
                  "algorithmic_gb": gb, "achieved_gbs": gb / (kern * 1e-3), "peak_gbs": peaks["hbm_gbs"],
                  "frac": gb / (kern * 1e-3) / peaks["hbm_gbs"], "torch_same_device_ms": theirs}))
