"""tf32x3 GEMM error vs K, and its bias (coherent part), against an fp64 product; torch fp32 beside it."""
import sys
sys.path.insert(0, ".")
import torch
from interactron_b200.ops import CudaOps
ops = CudaOps()
torch.manual_seed(0)
import os
print("ITN_GEMM_RZ_COMP =", os.environ.get("ITN_GEMM_RZ_COMP", "(default)"))
for K in (256, 2048, 4608):
    for kind in ("randn", "positive", "relu_x_w"):
        a = torch.randn(1024, K, device="cuda")
        b = torch.randn(512, K, device="cuda")
        if kind == "positive":
            a, b = a.abs(), b.abs()
        if kind == "relu_x_w":
            a = a.relu()
        ref = a.double() @ b.double().t()
        out = ops.matmul(a, b.t())
        t32 = a @ b.t()
        e = (out.double() - ref)
        rel = (e.norm() / ref.norm()).item()
        bias = (e.mean() / ref.abs().mean()).item()
        relt = ((t32.double() - ref).norm() / ref.norm()).item()
        print(f"K={K:5d} {kind:9s} tf32x3 rel {rel:.2e} mean-bias {bias:+.2e} | torch fp32 rel {relt:.2e}")
