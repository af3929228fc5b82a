"""Fused attention kernels (itn_attention_fwd / itn_attention_bwd, through the C ABI) against a float64
torch restatement of what the reference computes at models/gpt.py:43-53 and inside nn.MultiheadAttention
(models/detr_models/transformer.py:154-155,219-226): softmax(scale q k^T + key_padding_mask) v and its
autograd gradients.  Tolerance 2e-5 relative L2 (tf32x3 arithmetic; measured 1.3e-6 ... 4e-6), on the five
shape classes of the path, for both kernel families (sequential-phase and software-pipelined) and against the
unfused GEMM -> softmax -> GEMM chain the CPU simulation still uses.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = 2e-5


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def ops():
    from interactron_b200.ops import CudaOps
    return CudaOps()


def reference(q, k, v, nh, scale, kmask, dO):
    B, Lq, D = q.shape
    Lk, hd = k.shape[1], D // nh
    q64, k64, v64 = (t.detach().double().requires_grad_(True) for t in (q, k, v))
    qh = q64.view(B, Lq, nh, hd).permute(0, 2, 1, 3)
    kh = k64.view(B, Lk, nh, hd).permute(0, 2, 1, 3)
    vh = v64.view(B, Lk, nh, hd).permute(0, 2, 1, 3)
    s = scale * qh @ kh.transpose(-1, -2)
    if kmask is not None:
        s = s.masked_fill(kmask.bool()[:, None, None, :], float("-inf"))
    lse2 = torch.logsumexp(s, -1) * 1.4426950408889634
    o = (torch.softmax(s, -1) @ vh).permute(0, 2, 1, 3).reshape(B, Lq, D)
    o.backward(dO.double())
    return o.detach(), lse2, q64.grad, k64.grad, v64.grad


def make_inputs(ops, B, Lq, Lk, nh, hd, selfattn, masked, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    D, dev = nh * hd, ops.device
    if selfattn:                                   # q and k are column halves of one buffer, as on the path
        qk = torch.randn(B, Lq, 2 * D, device=dev, generator=g)
        q, k = qk[..., :D], qk[..., D:]
    else:
        q = 2.0 * torch.randn(B, Lq, D, device=dev, generator=g)
        k = torch.randn(B, Lk, D, device=dev, generator=g)
    v = torch.randn(B, Lk, D, device=dev, generator=g)
    dO = torch.randn(B, Lq, D, device=dev, generator=g)
    kmask = None
    if masked:
        kmask = torch.zeros(B, Lk, dtype=torch.uint8, device=dev)
        for b in range(B):
            kmask[b, Lk - 3 - 7 * b:] = 1
            kmask[b, 5] = 1
    return q, k, v, dO, kmask


def run_fused(ops, q, k, v, dO, kmask, nh, scale, selfattn):
    B, Lq, D = q.shape
    Lk = k.shape[1]
    o, lse = ops.attention_fwd(q, k, v, nh, scale, kmask)
    if selfattn:
        dqk = torch.zeros(B, Lq, 2 * D, device=q.device)
        dq, dk = dqk[..., :D], dqk[..., D:]
    else:
        dq, dk = torch.zeros(B, Lq, D, device=q.device), torch.zeros(B, Lk, D, device=q.device)
    dv = torch.zeros(B, Lk, D, device=q.device)
    ops.attention_bwd(dO, q, k, v, o, lse, nh, scale, kmask, dq, dk, dv)
    torch.cuda.synchronize()
    return o, lse, dq, dk, dv


SHAPES = [
    # B, Lq, Lk, nh, hd, self-attention, key mask
    (2, 361, 361, 8, 32, True, False),      # DETR encoder self-attention (transformer.py:154-155)
    (3, 361, 361, 8, 32, True, True),       #   ... with padded keys
    (2, 50, 50, 8, 32, True, False),        # DETR decoder self-attention (transformer.py:219)
    (2, 50, 361, 8, 32, False, True),       # DETR decoder cross-attention (transformer.py:222-226)
    (2, 255, 255, 8, 64, True, False),      # fusion B self-attention (new_transformer.py:23-25)
    (2, 255, 1805, 8, 64, False, False),    # fusion B cross-attention
    (1, 2060, 2060, 8, 64, False, False),   # fusion A, GPT full attention (gpt.py:43-53)
    (1, 416, 416, 8, 64, False, False),     # fusion A at the first policy step (interactron.py:174-197)
    (1, 1, 7, 1, 32, False, False),         # degenerate: one query, seven keys
    (1, 130, 129, 2, 64, False, True),      # ragged tiles on both sides
]


@pytest.mark.parametrize("family", ["default", "seq", "pipe"])
@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(str(v) for v in s))
def test_fused_attention_matches_float64(ops, shape, family, monkeypatch):
    B, Lq, Lk, nh, hd, selfattn, masked = shape
    if family != "default":
        for name in ("ITN_ATTN_FWD", "ITN_ATTN_DQ", "ITN_ATTN_DKV"):
            monkeypatch.setenv(name, family)
    q, k, v, dO, kmask = make_inputs(ops, B, Lq, Lk, nh, hd, selfattn, masked)
    scale = hd ** -0.5
    assert ops.attention_supported(q, k, v, nh)
    o_ref, lse_ref, dq_ref, dk_ref, dv_ref = reference(q, k, v, nh, scale, kmask, dO)
    o, lse, dq, dk, dv = run_fused(ops, q, k, v, dO, kmask, nh, scale, selfattn)
    for name, a, b in (("o", o, o_ref), ("lse", lse, lse_ref), ("dq", dq, dq_ref), ("dk", dk, dk_ref),
                       ("dv", dv, dv_ref)):
        assert torch.isfinite(a).all(), name
        assert rel(a, b) < TOL, (name, rel(a, b))
    if masked:                                      # padded keys receive no gradient
        m = kmask.bool()
        assert dk[m].abs().max().item() == 0.0 and dv[m].abs().max().item() == 0.0


def test_fused_equals_unfused_chain(ops):
    """Same inputs through the unfused kernels (QK^T GEMM, softmax, PV GEMM and their backward)."""
    from interactron_b200 import layers
    B, Lq, Lk, nh, hd = 4, 361, 361, 8, 32
    q, k, v, dO, kmask = make_inputs(ops, B, Lq, Lk, nh, hd, True, True, seed=3)
    scale = hd ** -0.5
    o, lse, dq, dk, dv = run_fused(ops, q, k, v, dO, kmask, nh, scale, True)
    ou, P = layers._attention_fwd_unfused(ops, q, k, v, B, Lq, Lk, nh, hd, scale, kmask)
    D = nh * hd
    dqk = torch.zeros(B, Lq, 2 * D, device=q.device)
    dvu = torch.zeros(B, Lk, D, device=q.device)
    layers.attention_bwd(ops, dO, q, k, v, P, B, Lq, Lk, nh, hd, scale, dqk[..., :D], dqk[..., D:], dvu)
    torch.cuda.synchronize()
    assert rel(o, ou) < TOL and rel(dq, dqk[..., :D]) < TOL and rel(dk, dqk[..., D:]) < TOL and rel(dv, dvu) < TOL


def test_fused_attention_is_bit_reproducible(ops):
    q, k, v, dO, kmask = make_inputs(ops, 2, 361, 361, 8, 32, True, True, seed=5)
    a = run_fused(ops, q, k, v, dO, kmask, 8, 32 ** -0.5, True)
    b = run_fused(ops, q, k, v, dO, kmask, 8, 32 ** -0.5, True)
    for x, y in zip(a, b):
        assert torch.equal(x, y)


def test_unsupported_views_are_refused(ops):
    from interactron_b200 import _lib
    q = torch.randn(1, 16, 48, device=ops.device)           # head dim 48 is not built
    assert not ops.attention_supported(q, q, q, 1)
    with pytest.raises(_lib.ItnError):
        ops.attention_fwd(q, q, q, 1, 1.0, None)
