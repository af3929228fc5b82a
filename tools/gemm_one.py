"""Launch one GEMM shape a few times (for ncu): python tools/gemm_one.py M N K [prec] [bn] [batch]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interactron_b200.ops import CudaOps  # noqa: E402

M, N, K = (int(x) for x in sys.argv[1:4])
prec = sys.argv[4] if len(sys.argv) > 4 else "tf32x3"
if len(sys.argv) > 5:
    os.environ["ITN_GEMM_BN"] = sys.argv[5]
batch = (int(sys.argv[6]),) if len(sys.argv) > 6 else ()
ops = CudaOps()
ops.precision = prec
a = torch.randn(*batch, M, K, device="cuda")
w = torch.randn(*batch, N, K, device="cuda")
out = torch.empty(*batch, M, N, device="cuda")
for _ in range(5):
    ops.matmul(a, w.transpose(-1, -2), out=out)
torch.cuda.synchronize()
# 10 launches replayed from a CUDA graph, so short kernels are not bound by the Python launch path
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(10):
        ops.matmul(a, w.transpose(-1, -2), out=out)
g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
g.replay()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"M={M} N={N} K={K} batch={batch} {prec}: {ms*1e3:.1f} us  {2.0*M*N*K*max(1, *batch, 1)/ms/1e9:.1f} TF/s")
