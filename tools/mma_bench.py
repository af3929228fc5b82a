"""Cycles per tcgen05.mma.kind::tf32 (128 x N x 8) by operand source / layout: what bounds the attention kernels."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interactron_b200 import _lib  # noqa: E402

lib = _lib.load()
lib.itn_debug_mma_bench.restype = C.c_int
lib.itn_debug_mma_bench.argtypes = [C.c_int] * 5 + [C.c_void_p, C.c_void_p]
torch.zeros(1, device="cuda")
for grid in (1, 148):
    out = torch.zeros(2 * grid, dtype=torch.int64, device="cuda")
    for n in (32, 64, 128, 256):
        for a_tmem in (1, 0):
            for b_mn in (0, 1):
                res = []
                for iters in (64, 1024):
                    _lib.check(lib.itn_debug_mma_bench(n, a_tmem, b_mn, iters, grid, out.data_ptr(), None))
                    torch.cuda.synchronize()
                    o = out.view(grid, 2).double()
                    res.append((iters, o[:, 0].mean().item(), o[:, 1].mean().item()))
                (i0, t0, s0), (i1, t1, s1) = res
                print(f"grid={grid:3d} N={n:3d} A={'tmem' if a_tmem else 'smem'} B={'MN' if b_mn else 'K '}: "
                      f"{(t1 - t0) / (i1 - i0):6.1f} cycles/MMA (issue {(s1 - s0) / (i1 - i0):5.1f}), "
                      f"fixed {t0 - i0 * (t1 - t0) / (i1 - i0):7.0f} cycles", flush=True)
