"""Accuracy of the fused attention forward against the unfused GEMM -> softmax -> GEMM chain, both vs float64:
relative L2 error and the component of the error along the true output (a scale bias), per shape class and
input regime (flat softmax / sharp softmax, zero-mean / offset values).  python tools/attn_precision.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interactron_b200 import layers  # noqa: E402
from interactron_b200.ops import CudaOps  # noqa: E402


def ref(q, k, v, nh, scale):
    B, Lq, D = q.shape
    Lk, hd = k.shape[1], D // nh
    qh = q.double().view(B, Lq, nh, hd).permute(0, 2, 1, 3)
    kh = k.double().view(B, Lk, nh, hd).permute(0, 2, 1, 3)
    vh = v.double().view(B, Lk, nh, hd).permute(0, 2, 1, 3)
    p = torch.softmax(scale * qh @ kh.transpose(-1, -2), -1)
    return (p @ vh).permute(0, 2, 1, 3).reshape(B, Lq, D)


def ref_bwd(q, k, v, nh, scale, dO):
    q64, k64, v64 = (t.detach().double().requires_grad_(True) for t in (q, k, v))
    ref(q64, k64, v64, nh, scale).backward(dO.double())
    return q64.grad, k64.grad, v64.grad


def stats(o, o64):
    e = o.double() - o64
    return (e.norm() / o64.norm()).item(), ((e * o64).sum() / (o64 * o64).sum()).item()


ops = CudaOps(torch.device("cuda"))
g = torch.Generator(device="cuda").manual_seed(0)
for name, B, Lq, Lk, nh, hd in [("enc", 8, 361, 361, 8, 32), ("deccross", 8, 50, 361, 8, 32),
                                ("fusA", 2, 1805, 1805, 8, 64), ("fusB", 2, 250, 1805, 8, 64)]:
    D = nh * hd
    for regime, qs, voff in [("flat/zero-mean", 0.3, 0.0), ("flat/offset", 0.3, 3.0), ("sharp/zero-mean", 2.5, 0.0),
                             ("sharp/offset", 2.5, 3.0)]:
        q = torch.randn(B, Lq, D, device="cuda", generator=g) * qs
        k = torch.randn(B, Lk, D, device="cuda", generator=g)
        v = torch.randn(B, Lk, D, device="cuda", generator=g) + voff
        scale = hd ** -0.5
        o64 = ref(q, k, v, nh, scale)
        of, _ = ops.attention_fwd(q, k, v, nh, scale, None)
        ou, _ = layers._attention_fwd_unfused(ops, q, k, v, B, Lq, Lk, nh, hd, scale, None)
        o32 = ref(q.float(), k.float(), v.float(), nh, scale)  # same fp64 math (inputs are fp32 already)
        qh = q.view(B, Lq, nh, hd).permute(0, 2, 1, 3)
        kh = k.view(B, Lk, nh, hd).permute(0, 2, 1, 3)
        vh = v.view(B, Lk, nh, hd).permute(0, 2, 1, 3)
        torch.backends.cuda.matmul.allow_tf32 = False
        ot = (torch.softmax(scale * qh @ kh.transpose(-1, -2), -1) @ vh).permute(0, 2, 1, 3).reshape(B, Lq, D)
        rf, bf = stats(of, o64)
        ru, bu = stats(ou, o64)
        rt, bt = stats(ot, o64)
        print(f"{name:9s} {regime:16s} fused {rf:.2e} (bias {bf:+.2e})   unfused {ru:.2e} (bias {bu:+.2e})   "
              f"torch fp32 {rt:.2e} (bias {bt:+.2e})")
        if voff == 0.0:
            dO = torch.randn(B, Lq, D, device="cuda", generator=g)
            g64 = ref_bwd(q, k, v, nh, scale, dO)
            of, lse = ops.attention_fwd(q, k, v, nh, scale, None)
            dq, dk, dv = (torch.empty_like(t) for t in (q, k, v))
            ops.attention_bwd(dO, q, k, v, of, lse, nh, scale, None, dq, dk, dv)
            print("          backward fused: " + "  ".join(f"{n} {stats(a, b)[0]:.2e} (bias {stats(a, b)[1]:+.2e})"
                                                       for n, a, b in zip(("dq", "dk", "dv"), (dq, dk, dv), g64)))
