"""CTA-pair (cta_group::2) variant of the tf32x3 GEMM against the single-CTA kernel (must be bit-identical: same
products in the same order) and fp64, plus timing.  Child processes: the env switch is read once per process."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, torch
sys.path.insert(0, %r)
from interactron_b200.ops import CudaOps
ops = CudaOps()
torch.manual_seed(0)
out = {}
import json
SHAPES = json.loads(os.environ.get("SHAPES", "null")) or [(1000, 256, 96, False, False, False), (57760, 256, 2048, True, True, True), (16480, 2048, 512, False, False, False),
                             (57760, 2048, 256, True, False, True), (20000, 1236, 256, True, False, False), (33333, 512, 1496, False, False, False)]
for (M, N, K, bias, res, presplit) in SHAPES:
    a = torch.randn(M, K, device="cuda")
    w = torch.randn(N, K, device="cuda")
    if presplit:
        ops.register_presplit(w)
    b = torch.randn(N, device="cuda") if bias else None
    r = torch.randn(M, N, device="cuda") if res else None
    y = ops.matmul(a, w.t(), bias=b, residual=r, act="relu" if bias and not res else None)
    torch.cuda.synchronize()
    ref = a.double() @ w.double().t()
    if b is not None: ref = ref + b.double()
    if bias and not res: ref = ref.relu()
    if r is not None: ref = ref + r.double()
    err = ((y.double() - ref).norm() / ref.norm()).item()
    for _ in range(3): ops.matmul(a, w.t(), bias=b, residual=r, out=y)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): ops.matmul(a, w.t(), bias=b, residual=r, out=y)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    print(f"{M}x{N}x{K} bias={bias} res={res} presplit={presplit}: err {err:.2e}  {us:8.1f} us {2.0*M*N*K/us/1e6:7.1f} TF/s  sum {y.double().sum().item():.10e}", flush=True)
''' % ROOT

for pair in ("0", "1"):
    print(f"== ITN_GEMM_PAIR={pair}", flush=True)
    env = dict(os.environ, ITN_GEMM_PAIR=pair)
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=300)
    print(r.stdout[-3000:], r.stderr[-2000:], flush=True)
