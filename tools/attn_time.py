"""Default-dispatch timing of the fused attention kernels on the workload's shape classes (fwd, dQ, dK/dV)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interactron_b200.ops import CudaOps  # noqa: E402

SHAPES = {"enc": (160, 361, 361, 8, 32), "decself": (160, 50, 50, 8, 32), "deccross": (160, 50, 361, 8, 32),
          "fusBself": (32, 255, 255, 8, 64), "fusB": (32, 255, 1805, 8, 64), "fusA": (16, 2060, 2060, 8, 64)}


def timeit(fn, n=5):
    fn()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(n):
        fn()
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / n * 1e3


ops = CudaOps()
tot = 0.0
for name, (B, Lq, Lk, nh, hd) in SHAPES.items():
    D = nh * hd
    q, k, v, dO = (torch.randn(B, L, D, device="cuda") for L in (Lq, Lk, Lk, Lq))
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    scale = hd ** -0.5
    o, lse = ops.attention_fwd(q, k, v, nh, scale, None)
    f = timeit(lambda: ops.attention_fwd(q, k, v, nh, scale, None))
    os.environ["ITN_ATTN_BWD_ONLY"] = "1"
    a = timeit(lambda: ops.attention_bwd(dO, q, k, v, o, lse, nh, scale, None, dq, dk, dv))
    os.environ["ITN_ATTN_BWD_ONLY"] = "2"
    b = timeit(lambda: ops.attention_bwd(dO, q, k, v, o, lse, nh, scale, None, dq, dk, dv))
    os.environ.pop("ITN_ATTN_BWD_ONLY")
    tot += f + a + b
    print(f"{name:9s} fwd {f:8.1f}  dq {a:8.1f}  dkv {b:8.1f} us", flush=True)
print(f"sum {tot:.1f} us")
