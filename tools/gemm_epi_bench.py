"""Times one GEMM shape under different fused-epilogue options (diagnostic)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interactron_b200.ops import CudaOps  # noqa: E402

M, N, K = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (225000, 256, 64)
ops = CudaOps()
a = torch.randn(M, K, device="cuda")
w = torch.randn(N, K, device="cuda")
bias = torch.randn(N, device="cuda")
res = torch.randn(M, N, device="cuda")
out = torch.empty(M, N, device="cuda")
pre = torch.empty(M, N, device="cuda")


def bench(name, **kw):
    for _ in range(3):
        ops.matmul(a, w.t(), out=out, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.matmul(a, w.t(), out=out, **kw)
    e1.record()
    torch.cuda.synchronize()
    print(f"  {name:34s} {e0.elapsed_time(e1)/10*1e3:9.1f} us", flush=True)


for bn in ("0", "64", "128", "256"):
    os.environ["ITN_GEMM_BN"] = bn
    print(f"M={M} N={N} K={K} BN override {bn}")
    bench("plain")
    bench("bias")
    bench("bias+relu", bias=bias, act="relu")
    bench("residual", residual=res)
    bench("bias+relu_after+residual", bias=bias, act="relu", residual=res, act_after_residual=True)
    bench("relu_mask(aux)", epi="relu_mask", aux=res)
    bench("accumulate", accumulate=True)
    bench("out_pre", out_pre=pre)
