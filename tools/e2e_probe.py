"""Where the end-to-end gap of predict() on host buffers goes: host-side mask sampling, the H2D copy, the step,
the output clones / read-back; copy-first path vs graph.PipelinedPredict.  python tools/e2e_probe.py [E]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import interactron_b200 as ib  # noqa: E402
from interactron_b200.episode import sample_masks_host  # noqa: E402
from interactron_b200.synthetic import collate_episodes, synthetic_episode  # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 62
model = ib.build_model(ib.default_config("interactron_random", weights="synthetic").MODEL).cuda().eval()
data = collate_episodes([synthetic_episode(i, with_targets=False) for i in range(E)])
data["frames"] = data["frames"].pin_memory()
data["masks"] = data["masks"].pin_memory()
t0 = time.perf_counter()
for _ in range(5):
    m = sample_masks_host(data["masks"])
print(f"sample_masks_host: {(time.perf_counter() - t0) / 5 * 1e3:.2f} ms (host)")
dev = torch.empty(data["frames"].shape, device="cuda")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    dev.copy_(data["frames"], non_blocking=True)
e1.record()
torch.cuda.synchronize()
print(f"H2D of the frames ({dev.numel() * 4 / 1e6:.0f} MB): {e0.elapsed_time(e1) / 5:.2f} ms = {dev.numel() * 4 / 1e6 / (e0.elapsed_time(e1) / 5):.1f} GB/s")
for piped in (False, True, False, True):
    model.pipelined_input = piped
    for _ in range(3):
        model.predict(data)
    torch.cuda.synchronize()
    ts, evs = [], []
    for _ in range(6):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record()
        o = model.predict(data)
        lg = o["pred_logits"].cpu()
        b.record()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
        evs.append(a.elapsed_time(b))
    print(f"pipelined_input={piped}: predict + read-back {sorted(evs)[len(evs) // 2]:.2f} ms (events), {sorted(ts)[len(ts) // 2]:.2f} ms (host clock)")
