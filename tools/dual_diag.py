"""Compares every cached intermediate (primal and tangent) of a dual decoder layer: CUDA vs fp64 sim."""
import sys
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import torch
from interactron_b200.ops import CudaOps
from oracle.sim_ops import SimOps
import test_dual_gpu as T

ops, sim = CudaOps(), SimOps(torch.float64)
import interactron_b200.layers as layers
captured = {}
orig = layers.decoder_layer_fwd

def spy(o, *a, **k):
    out = orig(o, *a, **k)
    captured[o.o.name] = out
    return out

layers.decoder_layer_fwd = spy
try:
    T.test_dual_decoder_layer_matches_simulation(ops, sim)
except AssertionError as e:
    print("assert:", str(e)[:80])
cs, cc = captured["sim"][2], captured["cuda"][2]
for k in cs:
    a, b = cc[k], cs[k]
    if not hasattr(a, "p"):
        continue
    ep = T.rel(a.p, b.p) if a.p.dtype.is_floating_point else 0
    et = T.rel(a.t, b.t) if (a.t is not None and b.t is not None) else -1
    print(f"{k:10s} primal {ep:.2e} tangent {et:.2e}")
for i, nm in enumerate(("out", "out_r")):
    a, b = captured["cuda"][i], captured["sim"][i]
    print(nm, T.rel(a.p, b.p), T.rel(a.t, b.t))
