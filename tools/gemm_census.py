"""Per-shape GEMM census of one eager adapt+detect step: python tools/gemm_census.py [E] [workload]
Every ops.matmul is bracketed by a CUDA-event pair; a device-side sleep in front of the step lets the
host run ahead so the pairs measure execution, not launch gaps."""
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import interactron_b200 as ib  # noqa: E402
from interactron_b200.synthetic import collate_episodes, synthetic_episode  # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 8
name = sys.argv[2] if len(sys.argv) > 2 else "interactron_random"
model = ib.build_model(ib.default_config(name, weights="synthetic").MODEL).cuda().eval()
loop = model._get_loop()
d = collate_episodes([synthetic_episode(i, with_targets=False) for i in range(E)])
f, m = d["frames"].cuda(), d["masks"].cuda()
loop.adapt_detect(f, m, post_frames=(0,))
torch.cuda.synchronize()
ops = loop.ops
orig = ops.matmul
rec = []


def major(t):
    return "K" if t.stride(-1) == 1 else "MN"


def timed(a, b, **kw):
    M, K = a.shape[-2], a.shape[-1]
    N = b.shape[-1]
    nb = 1
    for x, y in zip(([1, 1] + list(a.shape[:-2]))[-2:], ([1, 1] + list(b.shape[:-2]))[-2:]):
        nb *= max(x, y)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = orig(a, b, **kw)
    e1.record()
    bm = "K" if b.stride(-2) == 1 else "MN"
    rec.append(((M, N, K, nb, major(a), bm), e0, e1))
    return out


ops.matmul = timed
torch.cuda._sleep(int(0.08 * 1.9e9))
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
loop.adapt_detect(f, m, post_frames=(0,))
t1.record()
torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for key, e0, e1 in rec:
    agg[key][0] += 1
    agg[key][1] += e0.elapsed_time(e1)
tot = sum(v[1] for v in agg.values())
fl = sum(2.0 * k[0] * k[1] * k[2] * k[3] * v[0] for k, v in agg.items())
print(f"E={E} {name}: {len(rec)} GEMMs, {tot:.2f} ms, {fl/1e9:.0f} GF, {fl/tot/1e9:.1f} TF/s; step {t0.elapsed_time(t1):.1f} ms")
print(f"{'M':>8} {'N':>6} {'K':>6} {'batch':>6} A/B  {'n':>4} {'ms':>8} {'us/launch':>10} {'TF/s':>7} {'share':>6}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    M, N, K, nb, am, bm = k
    print(f"{M:8d} {N:6d} {K:6d} {nb:6d} {am}/{bm:2s} {v[0]:4d} {v[1]:8.3f} {v[1]/v[0]*1e3:10.1f} "
          f"{2.0*M*N*K*nb*v[0]/v[1]/1e9:7.1f} {100*v[1]/tot:5.1f}%")
