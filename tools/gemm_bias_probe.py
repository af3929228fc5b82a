"""Scale bias of the tf32x3 GEMM (component of the error along the exact result) and its rel-L2 error against
float64, per operand regime, for the round-toward-zero compensation constant in ITN_GEMM_RZ_COMP (read once per
process: run once per value).  python tools/gemm_bias_probe.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interactron_b200.ops import CudaOps  # noqa: E402

ops = CudaOps()
g = torch.Generator(device="cuda").manual_seed(0)
print("ITN_GEMM_RZ_COMP =", os.environ.get("ITN_GEMM_RZ_COMP", "(default 0.59)"), " split =", ops.split_acc)
for K in (256, 512, 2048):
    for kind in ("randn x randn", "softmax-like x randn", "relu x randn", "randn+1 x randn+1"):
        a = torch.randn(2048, K, device="cuda", generator=g)
        b = torch.randn(512, K, device="cuda", generator=g)
        if kind.startswith("softmax"):
            a = torch.softmax(a * 2, -1)
        elif kind.startswith("relu"):
            a = torch.relu(a)
        elif kind.startswith("randn+1"):
            a, b = a + 1, b + 1
        want = a.double() @ b.double().t()
        y = ops.matmul(a, b.t()).double()
        e = y - want
        print(f"K={K:5d} {kind:22s} rel {float(e.norm() / want.norm()):.2e}  scale bias {float((e * want).sum() / (want * want).sum()):+.2e}")
