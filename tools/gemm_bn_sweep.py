"""Tile-width sweep for the small / few-tile GEMM shapes of the step: time per launch at forced BN (ITN_GEMM_BN) vs
the library's own choice.  python tools/gemm_bn_sweep.py
Eager launches: below ~35 us per launch the figure is the HOST launch rate (tensor-map encoding + ctypes), not the kernel -
the step replays a CUDA graph; use tools/gemm_census.py or an ncu launch list for the small shapes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interactron_b200.ops import CudaOps  # noqa: E402

ops = CudaOps()
g = torch.Generator(device="cuda").manual_seed(0)


def t(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


SHAPES = [(62, 250, 256, 256), (62, 250, 256, 2048), (62, 250, 2048, 256), (62, 50, 256, 256), (62, 50, 256, 2048),
          (62, 50, 2048, 256), (1, 15500, 256, 256), (1, 15500, 256, 512), (1, 15810, 512, 512), (62, 255, 512, 512),
          (62, 361, 256, 2048), (1, 22382, 512, 256), (1, 3100, 256, 256), (62, 1805, 256, 256)]
for (B, M, N, K) in SHAPES:
    a = torch.randn(B, M, K, device="cuda", generator=g)
    w = torch.randn(B, N, K, device="cuda", generator=g)
    bias = torch.randn(B, N, device="cuda", generator=g)
    out = torch.empty(B, M, N, device="cuda")
    res = {}
    for bn in ("auto", "64", "128", "256"):
        if bn == "auto":
            os.environ.pop("ITN_GEMM_BN", None)
        else:
            os.environ["ITN_GEMM_BN"] = bn
        res[bn] = t(lambda: ops.matmul(a, w.transpose(-1, -2), bias=bias, out=out))
    os.environ.pop("ITN_GEMM_BN", None)
    gf = 2.0 * B * M * N * K
    print(f"{B:3d} x {M:6d} x {N:5d} x {K:5d}: " + "  ".join(f"BN={k:>4s} {v:7.1f} us ({gf / v / 1e6:5.1f} TF/s)" for k, v in res.items()), flush=True)
