"""Fused attention kernels (itn_attention_fwd / itn_attention_bwd) against a float64 torch reference on the
GPU, shape class by shape class, with localisation of the error (which head / row tile / column block) so one
run says where a layout bug sits.  Also times fused vs the unfused GEMM -> softmax -> GEMM chain.

    python tools/attn_check.py            # correctness on small batches + timing at the bench's sizes
    python tools/attn_check.py quick      # correctness only
"""
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interactron_b200 import layers  # noqa: E402
from interactron_b200.ops import CudaOps  # noqa: E402


def ref_attention(q, k, v, nh, scale, kmask, dO):
    B, Lq, D = q.shape
    Lk = k.shape[1]
    hd = D // nh
    q64, k64, v64 = (t.detach().double().requires_grad_(True) for t in (q, k, v))
    qh = q64.view(B, Lq, nh, hd).permute(0, 2, 1, 3)
    kh = k64.view(B, Lk, nh, hd).permute(0, 2, 1, 3)
    vh = v64.view(B, Lk, nh, hd).permute(0, 2, 1, 3)
    s = scale * qh @ kh.transpose(-1, -2)
    if kmask is not None:
        s = s.masked_fill(kmask.bool()[:, None, None, :], float("-inf"))
    lse2 = torch.logsumexp(s, -1) * 1.4426950408889634
    p = torch.softmax(s, -1)
    o = (p @ vh).permute(0, 2, 1, 3).reshape(B, Lq, D)
    o.backward(dO.double())
    return o.detach(), lse2.detach(), q64.grad, k64.grad, v64.grad


def rel(a, b):
    return ((a.double() - b).norm() / b.norm().clamp_min(1e-300)).item()


def localise(name, a, b, nh):
    """a, b [B, L, nh*hd]: print the worst head / 128-row tile / 32-row block."""
    B, L, D = a.shape
    hd = D // nh
    e = (a.double() - b).view(B, L, nh, hd)
    r = b.view(B, L, nh, hd)
    per_head = e.pow(2).sum((0, 1, 3)).sqrt() / r.pow(2).sum((0, 1, 3)).sqrt().clamp_min(1e-300)
    per_b = e.pow(2).sum((1, 2, 3)).sqrt() / r.pow(2).sum((1, 2, 3)).sqrt().clamp_min(1e-300)
    rows = e.pow(2).sum((0, 2, 3)).sqrt() / r.pow(2).sum((0, 2, 3)).sqrt().clamp_min(1e-300)
    per_d = e.pow(2).sum((0, 1, 2)).sqrt() / r.pow(2).sum((0, 1, 2)).sqrt().clamp_min(1e-300)
    blocks = [rows[i:i + 32].max().item() for i in range(0, L, 32)]
    print(f"    {name}: per-head {['%.1e' % x for x in per_head.tolist()]}")
    print(f"    {name}: per-batch {['%.1e' % x for x in per_b.tolist()[:8]]}")
    print(f"    {name}: per 32-row block (max row err) {['%.1e' % x for x in blocks[:24]]}")
    print(f"    {name}: per head-dim col {['%.1e' % x for x in per_d.tolist()]}")
    print(f"    {name}: nan {torch.isnan(a).sum().item()} inf {torch.isinf(a).sum().item()}")


def check(ops, B, Lq, Lk, nh, hd, masked=False, selfattn=False, seed=0, tol=2e-5):
    g = torch.Generator(device="cuda").manual_seed(seed)
    D = nh * hd
    dev = ops.device
    if selfattn:
        qk = torch.randn(B, Lq, 2 * D, device=dev, generator=g)
        q, k = qk[..., :D], qk[..., D:]
    else:
        q = torch.randn(B, Lq, D, device=dev, generator=g)
        k = torch.randn(B, Lk, D, device=dev, generator=g)
    # sharpen the softmax a little so it is not uniform
    q = q * 2.0 if not selfattn else q
    v = torch.randn(B, Lk, D, device=dev, generator=g)
    dO = torch.randn(B, Lq, D, device=dev, generator=g)
    kmask = None
    if masked:
        kmask = torch.zeros(B, Lk, dtype=torch.uint8, device=dev)
        for b in range(B):
            kmask[b, Lk - 1 - 7 * b - 3:] = 1
            kmask[b, 5] = 1
    scale = hd ** -0.5
    o_ref, lse_ref, dq_ref, dk_ref, dv_ref = ref_attention(q, k, v, nh, scale, kmask, dO)
    assert ops.attention_supported(q, k, v, nh)
    o, lse = ops.attention_fwd(q, k, v, nh, scale, kmask)
    dqk = torch.zeros(B, Lq if selfattn else 1, 2 * D, device=dev)
    if selfattn:
        dq, dk = dqk[..., :D], dqk[..., D:]
    else:
        dq, dk = torch.zeros(B, Lq, D, device=dev), torch.zeros(B, Lk, D, device=dev)
    dv = torch.zeros(B, Lk, D, device=dev)
    ops.attention_bwd(dO, q, k, v, o, lse, nh, scale, kmask, dq, dk, dv)
    torch.cuda.synchronize()
    errs = dict(o=rel(o, o_ref), lse=rel(lse, lse_ref), dq=rel(dq, dq_ref), dk=rel(dk, dk_ref), dv=rel(dv, dv_ref))
    ok = all(e < tol for e in errs.values())
    print(f"[{'ok' if ok else 'FAIL'}] B={B} Lq={Lq} Lk={Lk} nh={nh} hd={hd} mask={masked} self={selfattn}: "
          + " ".join(f"{k_}={e:.2e}" for k_, e in errs.items()))
    if not ok:
        for name, a, b_ in (("o", o, o_ref), ("dq", dq, dq_ref), ("dk", dk, dk_ref), ("dv", dv, dv_ref)):
            if errs[name] >= tol:
                localise(name, a.contiguous(), b_, nh)
        if errs["lse"] >= tol:
            e = (lse.double() - lse_ref).abs()
            print(f"    lse: max abs err {e.max().item():.3e}, first rows fused {lse[0, 0, :6].tolist()} ref {lse_ref[0, 0, :6].tolist()}")
    return ok


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(n):
        fn()
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / n * 1e3


def bench(ops, B, Lq, Lk, nh, hd, selfattn):
    D = nh * hd
    dev = ops.device
    if selfattn:
        qk = torch.randn(B, Lq, 2 * D, device=dev)
        q, k = qk[..., :D], qk[..., D:]
    else:
        q, k = torch.randn(B, Lq, D, device=dev), torch.randn(B, Lk, D, device=dev)
    v, dO = torch.randn(B, Lk, D, device=dev), torch.randn(B, Lq, D, device=dev)
    scale = hd ** -0.5
    dq, dk, dv = torch.empty_like(q.contiguous()), torch.empty_like(k.contiguous()), torch.empty_like(v)
    o, lse = ops.attention_fwd(q, k, v, nh, scale, None)
    t_f = timeit(lambda: ops.attention_fwd(q, k, v, nh, scale, None))
    t_b = timeit(lambda: ops.attention_bwd(dO, q, k, v, o, lse, nh, scale, None, dq, dk, dv))
    ou, P = layers._attention_fwd_unfused(ops, q, k, v, B, Lq, Lk, nh, hd, scale, None)
    t_fu = timeit(lambda: layers._attention_fwd_unfused(ops, q, k, v, B, Lq, Lk, nh, hd, scale, None))
    t_bu = timeit(lambda: layers.attention_bwd(ops, dO, q, k, v, P, B, Lq, Lk, nh, hd, scale, dq, dk, dv))
    fl = 4.0 * B * nh * Lq * Lk * hd
    print(f"time B={B} Lq={Lq} Lk={Lk} hd={hd}: fused fwd {t_f:8.1f} us ({fl / t_f * 1e-6:6.1f} TF/s) bwd {t_b:8.1f} us "
          f"({2.5 * fl / t_b * 1e-6:6.1f} TF/s) | unfused fwd {t_fu:8.1f} us bwd {t_bu:8.1f} us | "
          f"speed-up fwd {t_fu / t_f:.2f}x bwd {t_bu / t_b:.2f}x", flush=True)


def main():
    ops = CudaOps()
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    ok = True
    # smallest cases first: one tile, one block
    ok &= check(ops, 1, 50, 50, 8, 32, selfattn=True)
    ok &= check(ops, 2, 128, 64, 2, 32)
    ok &= check(ops, 2, 361, 361, 8, 32, selfattn=True)
    ok &= check(ops, 3, 361, 361, 8, 32, masked=True, selfattn=True)
    ok &= check(ops, 2, 50, 361, 8, 32, masked=True)
    ok &= check(ops, 2, 255, 255, 8, 64, selfattn=True)
    ok &= check(ops, 2, 255, 1805, 8, 64)
    ok &= check(ops, 1, 2060, 2060, 8, 64)
    ok &= check(ops, 1, 416, 416, 8, 64)
    print("ALL OK" if ok else "SOME FAILED", flush=True)
    if quick:
        return 0 if ok else 1
    bench(ops, 160, 361, 361, 8, 32, True)
    bench(ops, 160, 50, 50, 8, 32, True)
    bench(ops, 160, 50, 361, 8, 32, False)
    bench(ops, 32, 255, 255, 8, 64, True)
    bench(ops, 32, 255, 1805, 8, 64, False)
    bench(ops, 16, 2060, 2060, 8, 64, False)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
