"""train()-mode dropout on the GPU: the mask function of csrc/itn_philox.cuh against its numpy restatement
(oracle/philox.py) bit for bit, the fused attention kernels with dropout against a float64 reference that uses
the same mask, and train()-mode `predict()` against the float32 torch simulation of the kernels with the same
seed (the simulation itself is pinned to the unmodified reference with injected masks in tests/test_dropout_cpu.py)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    return ((a.double().cpu() - b.double().cpu()).norm() / b.double().cpu().norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def ops():
    from interactron_b200.ops import CudaOps
    return CudaOps()


def seed_tensor(v):
    return torch.tensor([v], dtype=torch.int64, device="cuda")


@pytest.mark.parametrize("rows,cols", [(7, 4), (1805, 256), (250, 2048), (1000, 361), (3, 1237)])
def test_dropout_kernel_mask_is_the_oracle_mask(ops, rows, cols):
    from oracle import philox
    seed, site, p = 0x1234_5678_9ABC_DEF, 77, 0.1
    x = torch.randn(rows, cols, device="cuda")
    key = (p, seed_tensor(seed), site)
    y = ops.dropout(x, key)
    keep = torch.from_numpy(philox.keep_mask(seed, site, rows, cols, p)).cuda()
    want = x * keep * (1.0 / (1.0 - 0.1))
    assert torch.equal(y != 0, keep & (x != 0))
    assert rel(y, want) < 1e-6
    res = torch.randn(rows, cols, device="cuda")
    assert rel(ops.dropout(x, key, residual=res), res + want) < 1e-6
    z = x.clone()
    ops.dropout(z, key, out=z)                                       # in place
    assert torch.equal(z, y)
    # a strided 2-D view (the padded probability rows of the unfused attention path)
    ld = (cols + 3) // 4 * 4 + 4
    buf = torch.zeros(rows, ld, device="cuda")
    buf[:, :cols] = x
    out = torch.zeros(rows, ld, device="cuda")
    ops.dropout(buf[:, :cols], key, out=out[:, :cols])
    assert torch.equal(out[:, :cols], y) and out[:, cols:].abs().max().item() == 0.0
    # another site / seed gives another mask
    assert not torch.equal(ops.dropout(x, (p, seed_tensor(seed), site + 1)) != 0, y != 0)
    assert not torch.equal(ops.dropout(x, (p, seed_tensor(seed + 1), site)) != 0, y != 0)


ATTN_SHAPES = [
    # B, Lq, Lk, nh, hd, key mask
    (2, 361, 361, 8, 32, True),        # pipelined forward / dQ / dK-dV kernels, hd 32
    (2, 50, 361, 8, 32, False),
    (2, 255, 1805, 8, 64, False),      # hd 64: pipelined forward / dQ, sequential dK-dV kernel
    (1, 416, 416, 8, 64, False),
]


@pytest.mark.parametrize("shape", ATTN_SHAPES, ids=lambda s: "x".join(str(v) for v in s))
def test_fused_attention_with_dropout_matches_float64(ops, shape):
    from oracle import philox
    B, Lq, Lk, nh, hd, masked = shape
    D = nh * hd
    g = torch.Generator(device="cuda").manual_seed(1)
    q = 2.0 * torch.randn(B, Lq, D, device="cuda", generator=g)
    k, v = torch.randn(B, Lk, D, device="cuda", generator=g), torch.randn(B, Lk, D, device="cuda", generator=g)
    dO = torch.randn(B, Lq, D, device="cuda", generator=g)
    kmask = None
    if masked:
        kmask = torch.zeros(B, Lk, dtype=torch.uint8, device="cuda")
        kmask[:, Lk - 9:] = 1
    seed, site, p = 987654321987, 2051, 0.1
    key = (p, seed_tensor(seed), site)
    scale = hd ** -0.5
    keep = torch.from_numpy(philox.keep_mask(seed, site, B * nh * Lq, Lk, p)).cuda().view(B, nh, Lq, Lk)
    q64, k64, v64 = (t.double().requires_grad_(True) for t in (q, k, v))
    s = scale * q64.view(B, Lq, nh, hd).permute(0, 2, 1, 3) @ k64.view(B, Lk, nh, hd).permute(0, 2, 3, 1)
    if masked:
        s = s.masked_fill(kmask.bool()[:, None, None, :], float("-inf"))
    pd = torch.softmax(s, -1) * keep * (1.0 / (1.0 - p))
    o_ref = (pd @ v64.view(B, Lk, nh, hd).permute(0, 2, 1, 3)).permute(0, 2, 1, 3).reshape(B, Lq, D)
    o_ref.backward(dO.double())
    o, lse = ops.attention_fwd(q, k, v, nh, scale, kmask, drop=key)
    dq, dk, dv = torch.zeros_like(q), torch.zeros_like(k), torch.zeros_like(v)
    ops.attention_bwd(dO, q, k, v, o, lse, nh, scale, kmask, dq, dk, dv, drop=key)
    torch.cuda.synchronize()
    for name, a, b in (("o", o, o_ref.detach()), ("dq", dq, q64.grad), ("dk", dk, k64.grad), ("dv", dv, v64.grad)):
        assert torch.isfinite(a).all(), name
        assert rel(a, b) < 2e-5, (name, rel(a, b))
    # the unfused chain (what the dual-number pass runs) with the same key agrees too
    from interactron_b200 import layers
    ou, ctx = layers._attention_fwd_unfused(ops, q, k, v, B, Lq, Lk, nh, hd, scale, kmask, drop=key)
    dq2, dk2, dv2 = torch.zeros_like(q), torch.zeros_like(k), torch.zeros_like(v)
    layers.attention_bwd(ops, dO, q, k, v, ctx, B, Lq, Lk, nh, hd, scale, dq2, dk2, dv2)
    for name, a, b in (("o", ou, o), ("dq", dq2, dq), ("dk", dk2, dk), ("dv", dv2, dv)):
        assert rel(a, b) < 2e-5, ("unfused " + name, rel(a, b))


@pytest.mark.parametrize("model_type", ["interactron_random", "interactron"])
def test_train_mode_predict_matches_the_simulation(model_type):
    import copy
    import interactron_b200 as ib
    from interactron_b200.synthetic import synthetic_episode
    from oracle.sim_ops import SimOps
    model = ib.build_model(ib.default_config(model_type, weights="synthetic").MODEL).cuda().train()
    sim_model = copy.deepcopy(model).cpu().train()
    sim_model._ops, sim_model._loop, sim_model._graphs = SimOps(), None, {}
    data = synthetic_episode(2)
    seed = 0x7777_1234_5678
    loop, sloop = model._get_loop(), sim_model._get_loop()
    loop.drop_seed = seed_tensor(seed)
    sloop.drop_seed = torch.tensor([seed], dtype=torch.int64)
    out = loop.adapt_detect(data["frames"].cuda(), data["masks"].cuda(), post_frames=(0,), train=True)
    want = sloop.adapt_detect(data["frames"], data["masks"], post_frames=(0,), train=True)
    assert rel(out["learned_loss"], want["learned_loss"]) < 1e-3
    assert rel(out["pred_logits"], want["pred_logits"]) < 1e-3
    assert rel(out["pred_boxes"], want["pred_boxes"]) < 1e-3
    ev = loop.adapt_detect(data["frames"].cuda(), data["masks"].cuda(), post_frames=(0,))
    assert rel(ev["pred_logits"], want["pred_logits"]) > 1e-2         # dropout is really applied


def test_train_mode_public_api_draws_fresh_masks_under_cuda_graphs():
    import interactron_b200 as ib
    from interactron_b200.synthetic import synthetic_episode
    torch.manual_seed(1234)
    model = ib.build_model(ib.default_config("interactron_random", weights="synthetic").MODEL).cuda().train()
    assert model.use_cuda_graph
    data = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in synthetic_episode(1).items()}
    a = model.predict(data)["pred_logits"].clone()
    b = model.predict(data)["pred_logits"].clone()                    # graph replay with a new seed
    assert rel(a, b) > 1e-3
    model._drop_gen.manual_seed(torch.initial_seed() & 0x7FFFFFFFFFFFFFFF)   # rewind the seed stream
    c = model.predict(data)["pred_logits"]
    assert torch.equal(a, c)
    model.eval()
    e1, e2 = model.predict(data)["pred_logits"].clone(), model.predict(data)["pred_logits"]
    assert torch.equal(e1, e2) and rel(e1, a) > 1e-3
