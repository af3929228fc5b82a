"""Timing experiments on the fused attention kernels: block widths and phase knock-outs (ITN_ATTN_DBG).
Numbers under a non-zero DBG are timings of a kernel that computes garbage; they only say where the time goes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interactron_b200.ops import CudaOps  # noqa: E402

SHAPES = {
    "enc": (160, 361, 361, 8, 32),
    "deccross": (160, 50, 361, 8, 32),
    "fusB": (32, 255, 1805, 8, 64),
    "fusA": (16, 2060, 2060, 8, 64),
}


def timeit(fn, n=4):
    fn()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(n):
        fn()
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / n * 1e3


def run(label, fn, dbgs):
    row = []
    for dbg in dbgs:
        os.environ["ITN_ATTN_DBG"] = str(dbg)
        row.append(f"{dbg}:{timeit(fn):7.1f}")
    os.environ["ITN_ATTN_DBG"] = "0"
    print(label + "  " + "  ".join(row), flush=True)


def main():
    ops = CudaOps()
    names = sys.argv[1:] or list(SHAPES)
    for name in names:
        B, Lq, Lk, nh, hd = SHAPES[name]
        D = nh * hd
        dev = ops.device
        q, k, v, dO = (torch.randn(B, L, D, device=dev) for L in (Lq, Lk, Lk, Lq))
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        scale = hd ** -0.5
        o, lse = ops.attention_fwd(q, k, v, nh, scale, None)
        fwd = lambda: ops.attention_fwd(q, k, v, nh, scale, None)
        bwd = lambda: ops.attention_bwd(dO, q, k, v, o, lse, nh, scale, None, dq, dk, dv)
        for pipe in (1, 0):
            os.environ["ITN_ATTN_PIPE"] = str(pipe)
            if pipe:
                fwd_blks, dq_blks, dkv_blks = [32, 64], ([32, 64] if hd == 32 else [32]), [32]
                dbgs = [0, 2, 4, 8, 16, 14, 30]
            else:
                fwd_blks = [64] if hd == 32 else [32]
                dq_blks = [32] if hd == 32 else [64]
                dkv_blks = [64] if hd == 32 else [32]
                dbgs = [0]
            for blk in fwd_blks:
                os.environ["ITN_ATTN_FWD_BLK"] = str(blk)
                run(f"{name} pipe={pipe} fwd blk={blk:3d}", fwd, dbgs)
            os.environ.pop("ITN_ATTN_FWD_BLK")
            os.environ["ITN_ATTN_BWD_ONLY"] = "1"
            for blk in dq_blks:
                os.environ["ITN_ATTN_DQ_BLK"] = str(blk)
                run(f"{name} pipe={pipe} dq  blk={blk:3d}", bwd, dbgs)
            os.environ.pop("ITN_ATTN_DQ_BLK")
            os.environ["ITN_ATTN_BWD_ONLY"] = "2"
            for blk in dkv_blks:
                os.environ["ITN_ATTN_DKV_BLK"] = str(blk)
                run(f"{name} pipe={pipe} dkv blk={blk:3d}", bwd, dbgs)
            os.environ.pop("ITN_ATTN_DKV_BLK")
            os.environ.pop("ITN_ATTN_BWD_ONLY")
        os.environ.pop("ITN_ATTN_PIPE")


if __name__ == "__main__":
    main()
