"""Parity report of the meta-training step on the GPU against tests/golden/<model>_forward.pt:
sampled relative-L2 error of the meta-gradient per parameter group vs the fp64 reference, next to the
fp32 reference's own distance from fp64.  Usage: python tools/meta_parity.py [model ...]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import interactron_b200 as ib  # noqa: E402
from interactron_b200.synthetic import collate_episodes, synthetic_episode  # noqa: E402

stride_of = lambda n: max(1, n // 256) | 1


def group(name):
    if name.startswith("fusion."):
        return "phi (fusion)"
    return "psi (in_proj)" if "in_proj" in name else "theta (fast weights)"


for name in sys.argv[1:] or ("interactron_random", "interactron"):
    gold = torch.load(os.path.join(ROOT, "tests", "golden", f"{name}_forward.pt"))
    model = ib.build_model(ib.default_config(name, weights="synthetic").MODEL).cuda().eval()
    data = collate_episodes([synthetic_episode(e) for e in gold["episodes"]])
    p, l = model(data, ridx=list(gold["ridx"]))
    g32, g64 = gold["fp32"][0], gold["fp64"][0]
    print(f"== {name}: losses (ours | ref fp32 | ref fp64)")
    for k in l:
        print(f"   {k:28s} {float(l[k]):.6f} | {float(g32['losses'][k]):.6f} | {float(g64['losses'][k]):.6f}")
    acc = {}
    per = []
    for n, prm in model.named_parameters():
        e64 = g64["grads"][n]
        if e64 is None or prm.grad is None:
            continue
        mine = prm.grad.detach().double().cpu().reshape(-1)[::stride_of(prm.numel())]
        s64, s32 = e64["sample"].double(), g32["grads"][n]["sample"].double()
        a = acc.setdefault(group(n), [0.0, 0.0, 0.0])
        a[0] += float((mine - s64).pow(2).sum())
        a[1] += float((s32 - s64).pow(2).sum())
        a[2] += float(s64.pow(2).sum())
        if float(s64.norm()) > 1e-9:
            per.append((float((mine - s64).norm() / s64.norm()), float((s32 - s64).norm() / s64.norm()), n))
    print("   group                    ours-vs-fp64   ref32-vs-fp64")
    tot = [0.0, 0.0, 0.0]
    for k, a in acc.items():
        print(f"   {k:24s} {(a[0] / a[2]) ** 0.5:.3e}      {(a[1] / a[2]) ** 0.5:.3e}")
        tot = [x + y for x, y in zip(tot, a)]
    print(f"   {'all parameters':24s} {(tot[0] / tot[2]) ** 0.5:.3e}      {(tot[1] / tot[2]) ** 0.5:.3e}")
    per.sort(reverse=True)
    print("   worst tensors (ours, ref32, name):")
    for e in per[:8]:
        print(f"      {e[0]:.3e} {e[1]:.3e} {e[2]}")
    import statistics
    print("   median per-tensor: ours %.3e ref32 %.3e" % (statistics.median(x[0] for x in per), statistics.median(x[1] for x in per)))
