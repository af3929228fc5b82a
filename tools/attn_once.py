"""One launch of each fused attention kernel (forward, dQ, dK/dV) per shape class of the workload, for ncu captures:
ncu --set full --clock-control none -k regex:attn_ -c 18 python tools/attn_once.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interactron_b200.ops import CudaOps  # noqa: E402

SHAPES = {"enc": (160, 361, 361, 8, 32), "decself": (160, 50, 50, 8, 32), "deccross": (160, 50, 361, 8, 32),
          "fusBself": (32, 255, 255, 8, 64), "fusB": (32, 255, 1805, 8, 64), "fusA": (16, 2060, 2060, 8, 64)}
ops = CudaOps()
for name, (B, Lq, Lk, nh, hd) in SHAPES.items():
    D = nh * hd
    q, k, v, dO = (torch.randn(B, L, D, device="cuda") for L in (Lq, Lk, Lk, Lq))
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    o, lse = ops.attention_fwd(q, k, v, nh, hd ** -0.5, None)
    ops.attention_bwd(dO, q, k, v, o, lse, nh, hd ** -0.5, None, dq, dk, dv)
    torch.cuda.synchronize()
    print(name, "done", flush=True)
