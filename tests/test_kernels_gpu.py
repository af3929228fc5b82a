"""Per-kernel parity of the CUDA ops (through the C ABI) against plain torch fp32/fp64.

Tolerances: TF32 GEMMs 2e-3 relative L2 against an fp64 product (TF32 keeps 10 mantissa
bits); everything else is fp32 arithmetic and must agree to 1e-5 relative.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

TF32_TOL = 2e-3
FP32_TOL = 1e-5


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def ops():
    """Single-pass TF32 mode with TF32-clean operand discipline (the kernels' raw behaviour)."""
    from interactron_b200.ops import CudaOps
    o = CudaOps()
    o.precision = "tf32"
    return o


@pytest.fixture(scope="module")
def ops3():
    """Default error-compensated tf32x3 mode."""
    from interactron_b200.ops import CudaOps
    o = CudaOps()
    assert o.precision == "tf32x3"
    return o


X3_TOL = 3e-5


@pytest.mark.parametrize("a_mn", [False, True])
@pytest.mark.parametrize("b_mn", [False, True])
@pytest.mark.parametrize("shape", [(128, 128, 32), (1805, 256, 256), (256, 2048, 1805), (364, 32, 361),
                                   (512, 1496, 252), (2060, 512, 2048), (250, 4, 256), (250, 256, 4)])
def test_gemm_tf32x3_is_fp32_accurate(ops3, a_mn, b_mn, shape):
    M, N, K = shape
    gen = torch.Generator(device="cuda").manual_seed(M + N * 3 + K * 7)
    a = _mk((M, K), a_mn, gen)
    b = _mk((K, N), not b_mn, gen)
    out = ops3.matmul(a, b)
    assert rel(out, a.double() @ b.double()) < X3_TOL


@pytest.mark.parametrize("bn", ["32", "64", "128", "256"])
def test_gemm_tf32x3_tile_widths_and_batches(ops3, bn, monkeypatch):
    monkeypatch.setenv("ITN_GEMM_BN", bn)
    gen = torch.Generator(device="cuda").manual_seed(int(bn) + 1)
    a = torch.randn(3, 300, 200, generator=gen, device="cuda")
    w = torch.randn(3, 520, 200, generator=gen, device="cuda")
    bias = torch.randn(520, generator=gen, device="cuda")
    out = ops3.matmul(a, w.transpose(-1, -2), bias=bias, act="relu")
    assert rel(out, torch.relu(a.double() @ w.double().transpose(-1, -2) + bias.double())) < X3_TOL


def _mk(shape, transposed, gen):
    if transposed:
        t = torch.randn(*shape[:-2], shape[-1], shape[-2], generator=gen, device="cuda")
        return t.transpose(-1, -2)
    return torch.randn(*shape, generator=gen, device="cuda")


@pytest.mark.parametrize("a_mn", [False, True])
@pytest.mark.parametrize("b_mn", [False, True])
@pytest.mark.parametrize("shape", [(128, 128, 32), (1805, 256, 256), (250, 1236, 256), (256, 2048, 1805),
                                   (1236, 256, 250), (364, 32, 361), (512, 1496, 252), (2060, 512, 2048)])
def test_gemm_tf32_majors(ops, a_mn, b_mn, shape):
    M, N, K = shape
    gen = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    a = _mk((M, K), a_mn, gen)
    b = _mk((K, N), not b_mn, gen)
    n0 = ops.n_tf32
    out = ops.matmul(a, b)
    # TMA needs the non-contiguous stride of each operand to be a multiple of 4 floats
    aligned = ((M if a_mn else K) % 4 == 0) and ((N if b_mn else K) % 4 == 0)
    assert (ops.n_tf32 == n0 + 1) == aligned, "tcgen05 path must serve every TMA-aligned shape"
    assert rel(out, a.double() @ b.double()) < TF32_TOL


@pytest.mark.parametrize("bn", ["32", "64", "128", "256"])
def test_gemm_tf32_tile_widths(ops, bn, monkeypatch):
    monkeypatch.setenv("ITN_GEMM_BN", bn)
    gen = torch.Generator(device="cuda").manual_seed(int(bn))
    a = torch.randn(3, 300, 200, generator=gen, device="cuda")
    w = torch.randn(3, 520, 200, generator=gen, device="cuda")
    out = ops.matmul(a, w.transpose(-1, -2))
    assert rel(out, a.double() @ w.double().transpose(-1, -2)) < TF32_TOL


def test_gemm_unaligned_goes_simt(ops):
    gen = torch.Generator(device="cuda").manual_seed(5)
    a = torch.randn(250, 6, generator=gen, device="cuda")      # row stride 6 floats: not TMA-able
    w = torch.randn(6, 510, generator=gen, device="cuda")
    n0 = ops.n_simt
    out = ops.matmul(a, w)
    assert ops.n_simt == n0 + 1
    assert rel(out, a.double() @ w.double()) < FP32_TOL


def test_gemm_rank1_outer_product(ops):
    """N=1 data-grad of the loss decoder: [250,1] @ [1,512]."""
    gen = torch.Generator(device="cuda").manual_seed(5)
    a = torch.randn(250, 1, generator=gen, device="cuda")
    w = torch.randn(1, 512, generator=gen, device="cuda")
    assert rel(ops.matmul(a, w), a.double() @ w.double()) < TF32_TOL


def tf32_rn(x):
    """Reference round-to-nearest(-away) to 10 mantissa bits."""
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def test_tf32_clean_operands_make_gemm_unbiased(ops):
    """With operands rounded-to-nearest to TF32 at the producer the tensor core result equals
    the fp64 product of the rounded operands up to fp32 accumulation error."""
    gen = torch.Generator(device="cuda").manual_seed(8)
    a = torch.randn(512, 1024, generator=gen, device="cuda")
    w = torch.randn(768, 1024, generator=gen, device="cuda")
    ar, wr = ops.round_tf32(a), ops.round_tf32(w)
    assert torch.equal(ar, tf32_rn(a)) and torch.equal(wr, tf32_rn(w))
    out = ops.matmul(ar, wr.t())
    assert rel(out, ar.double() @ wr.double().t()) < 5e-6
    assert rel(out, a.double() @ w.double().t()) < 4e-4
    o2 = ops.matmul(a, w.t(), rnd=True)
    assert torch.equal(o2, tf32_rn(ops.matmul(a, w.t())))


def test_gemm_heads_views(ops):
    gen = torch.Generator(device="cuda").manual_seed(11)
    F_, L, H, hd = 5, 361, 8, 32
    q, k, v = (torch.randn(F_, L, H * hd, generator=gen, device="cuda") for _ in range(3))
    qh, kh, vh = (t.view(F_, L, H, hd).permute(0, 2, 1, 3) for t in (q, k, v))
    ldp = (L + 3) // 4 * 4
    s = torch.zeros(F_, H, L, ldp, device="cuda")[..., :L]
    ops.matmul(qh, kh.transpose(-1, -2), out=s)
    assert rel(s, qh.double() @ kh.double().transpose(-1, -2)) < TF32_TOL
    o = torch.empty(F_, L, H * hd, device="cuda")
    oh = o.view(F_, L, H, hd).permute(0, 2, 1, 3)
    ops.matmul(s, vh, out=oh)
    assert rel(oh, s.double() @ vh.double()) < TF32_TOL
    dv = ops.matmul(s.transpose(-1, -2), oh)
    assert rel(dv, s.double().transpose(-1, -2) @ oh.double()) < TF32_TOL


@pytest.mark.parametrize("simt", [False, True])
def test_gemm_epilogues(ops, simt):
    ops.force_simt = simt
    try:
        gen = torch.Generator(device="cuda").manual_seed(3)
        M, N, K = 300, 260, 96
        a = torch.randn(M, K, generator=gen, device="cuda")
        w = torch.randn(N, K, generator=gen, device="cuda")
        bias = torch.randn(N, generator=gen, device="cuda")
        res = torch.randn(M, N, generator=gen, device="cuda")
        aux = torch.randn(M, N, generator=gen, device="cuda")
        base = a.double() @ w.double().t()
        tol = FP32_TOL if simt else TF32_TOL
        assert rel(ops.matmul(a, w.t(), bias=bias), base + bias.double()) < tol
        o = ops.matmul(a, w.t(), bias=bias, act="relu", residual=res)
        assert rel(o, torch.relu(base + bias.double()) + res.double()) < tol
        pre = torch.empty(M, N, device="cuda")
        o = ops.matmul(a, w.t(), bias=bias, act="gelu", out_pre=pre)
        assert rel(o, torch.nn.functional.gelu(base + bias.double())) < tol
        assert rel(pre, base + bias.double()) < tol
        o = ops.matmul(a, w.t(), epi="relu_mask", aux=aux)
        assert rel(o, base * (aux > 0).double()) < tol
        x = aux.double().requires_grad_(True)
        torch.nn.functional.gelu(x).sum().backward()
        o = ops.matmul(a, w.t(), epi="gelu_grad", aux=aux)
        assert rel(o, base * x.grad) < tol
        c = res.clone()
        ops.matmul(a, w.t(), out=c, accumulate=True, alpha=0.5)
        assert rel(c, 0.5 * base + res.double()) < tol
    finally:
        ops.force_simt = False


@pytest.mark.parametrize("cols", [256, 512])
@pytest.mark.parametrize("groups", [1, 3])
def test_layernorm_fwd_bwd(ops, cols, groups):
    gen = torch.Generator(device="cuda").manual_seed(cols + groups)
    rows = 3 * 250
    x = torch.randn(rows, cols, generator=gen, device="cuda") * 2 + 0.5
    gamma = torch.randn(groups, cols, generator=gen, device="cuda")
    beta = torch.randn(groups, cols, generator=gen, device="cuda")
    dy = torch.randn(rows, cols, generator=gen, device="cuda")
    y, y_r, mean, rstd = ops.layernorm_fwd(x, gamma, beta)
    dg, db = torch.empty(groups, cols, device="cuda"), torch.empty(groups, cols, device="cuda")
    dx, dx_r = ops.layernorm_bwd(dy, x, mean, rstd, gamma, dgamma=dg, dbeta=db)
    assert torch.equal(y_r, tf32_rn(y)) and torch.equal(dx_r, tf32_rn(dx))
    xr = x.double().view(groups, -1, cols).requires_grad_(True)
    gr = gamma.double().requires_grad_(True)
    br = beta.double().requires_grad_(True)
    yr = torch.nn.functional.layer_norm(xr, (cols,)) * gr[:, None] + br[:, None]
    yr.backward(dy.double().view(groups, -1, cols))
    assert rel(y, yr.reshape(rows, cols)) < FP32_TOL
    assert rel(dx, xr.grad.reshape(rows, cols)) < FP32_TOL
    assert rel(dg, gr.grad) < FP32_TOL
    assert rel(db, br.grad) < FP32_TOL


@pytest.mark.parametrize("cols", [50, 255, 361, 1805, 2060])
def test_softmax_fwd_bwd(ops, cols):
    gen = torch.Generator(device="cuda").manual_seed(cols)
    ld = (cols + 3) // 4 * 4
    rows = 97
    s = torch.randn(rows, ld, generator=gen, device="cuda") * 3
    dp = torch.randn(rows, ld, generator=gen, device="cuda")
    scale = 0.176
    sr = (s[:, :cols].double() * scale).requires_grad_(True)
    pr = torch.softmax(sr, -1)
    pr.backward(dp[:, :cols].double())
    p = ops.softmax_(s.clone(), cols, scale)
    # outputs are stored TF32-rounded (<= 2^-11 relative per element)
    assert rel(p[:, :cols], pr) < 4e-4
    ds = ops.softmax_bwd_(p, dp.clone(), cols, scale)
    pd = p[:, :cols].double()
    ref_ds = scale * pd * (dp[:, :cols].double() - (pd * dp[:, :cols].double()).sum(-1, keepdim=True))
    assert rel(ds[:, :cols], ref_ds) < 4e-4


@pytest.mark.parametrize("Lq,Lk,hd", [(361, 361, 32), (50, 361, 32), (255, 1805, 64)])
def test_gemm_padded_score_rows_vector_store(ops3, Lq, Lk, hd):
    ops = ops3
    """Attention scores: N % 4 != 0 written with 128-bit stores into rows that own their padding
    (itn_gemm_desc_t.c_pad); must equal the scalar-store result and leave zeros in the pad."""
    gen = torch.Generator(device="cuda").manual_seed(Lq + Lk)
    q = torch.randn(2, 3, Lq, hd, generator=gen, device="cuda")
    k = torch.randn(2, 3, Lk, hd, generator=gen, device="cuda")
    ld = (Lk + 3) // 4 * 4
    P0 = torch.full((2, 3, Lq, ld), 7.0, device="cuda")
    P1 = torch.full((2, 3, Lq, ld), 7.0, device="cuda")
    ops.matmul(q, k.transpose(-1, -2), out=P0[..., :Lk], alpha=0.25)
    ops.matmul(q, k.transpose(-1, -2), out=P1[..., :Lk], alpha=0.25, out_pad=True)
    assert torch.equal(P0[..., :Lk], P1[..., :Lk])
    assert (P0[..., Lk:] == 7.0).all() and (P1[..., Lk:] == 0.0).all()
    assert rel(P1[..., :Lk], 0.25 * q.double() @ k.double().transpose(-1, -2)) < X3_TOL


def test_softmax_unpadded_rows_take_generic_path(ops):
    """ld = 361 (rows not 16-byte aligned): the register-resident kernels do not apply."""
    gen = torch.Generator(device="cuda").manual_seed(5)
    s = torch.randn(40, 361, generator=gen, device="cuda") * 2
    dp = torch.randn(40, 361, generator=gen, device="cuda")
    ref = torch.softmax(s.double() * 0.5, -1)
    p = ops.softmax_(s.clone(), 361, 0.5)
    assert rel(p, ref) < 4e-4
    ds = ops.softmax_bwd_(p, dp.clone(), 361, 0.5)
    pd = p.double()
    assert rel(ds, 0.5 * pd * (dp.double() - (pd * dp.double()).sum(-1, keepdim=True))) < 4e-4


def test_softmax_key_mask(ops):
    gen = torch.Generator(device="cuda").manual_seed(1)
    B, H, L, Lk = 2, 4, 50, 361
    ld = 364
    s = torch.randn(B, H, L, ld, generator=gen, device="cuda")
    mask = (torch.rand(B, Lk, generator=gen, device="cuda") < 0.3)
    ref = torch.softmax(s[..., :Lk].double().masked_fill(mask[:, None, None, :], float("-inf")), -1)
    p = ops.softmax_(s.clone(), Lk, 1.0, key_mask=mask.to(torch.uint8).contiguous(), rows_per_mask=H * L)
    assert rel(p[..., :Lk], ref) < 4e-4


def test_colsum_add_copy_sigmoid_norm(ops):
    gen = torch.Generator(device="cuda").manual_seed(2)
    x = torch.randn(3, 1805, 256, generator=gen, device="cuda")
    assert rel(ops.colsum(x), x.double().sum(1)) < FP32_TOL
    for shape in ((1, 7220, 256), (2, 4096, 512), (1, 2053, 64)):      # two-pass path (rows split) + prime rows
        big = torch.randn(*shape, generator=gen, device="cuda")
        out = torch.zeros(shape[0], shape[2] + 8, device="cuda")
        ops.colsum(big, out=out[:, :shape[2]])
        assert rel(out[:, :shape[2]], big.double().sum(1)) < FP32_TOL and out[:, shape[2]:].abs().sum() == 0
    pos = torch.randn(1805, 256, generator=gen, device="cuda")
    assert rel(ops.add(x, pos), x + pos) < 1e-7
    dst = torch.zeros(250, 1496, device="cuda")
    src = torch.randn(250, 256, generator=gen, device="cuda")
    ops.copy2d_(dst[:, 4:260], src)
    assert torch.equal(dst[:, 4:260], src) and dst[:, :4].abs().sum() == 0
    z = torch.randn(250, 4, generator=gen, device="cuda")
    y = ops.sigmoid(z)
    assert rel(y, torch.sigmoid(z.double())) < FP32_TOL
    dy = torch.randn(250, 4, generator=gen, device="cuda")
    assert rel(ops.sigmoid_bwd(dy, y), dy.double() * y.double() * (1 - y.double())) < FP32_TOL
    v = torch.randn(4, 250, generator=gen, device="cuda")
    loss, dv = ops.l2norm_fwd_bwd(v)
    assert rel(loss, v.double().norm(dim=1)) < FP32_TOL
    assert rel(dv, v.double() / v.double().norm(dim=1, keepdim=True)) < FP32_TOL


@pytest.mark.parametrize("groups", [1, 4])
def test_sgd_clip_update_bit_exact(ops, groups):
    """The fused step must equal p - clip(lr*g) of utils/meta_utils.py:135-142 bit for bit."""
    gen = torch.Generator(device="cuda").manual_seed(9)
    n = 14_798_296 if groups == 1 else 1_000_004
    theta = torch.randn(n, generator=gen, device="cuda")
    g = torch.randn(groups, n, generator=gen, device="cuda") * 20
    out, out_r, mask = ops.sgd_clip_update(theta, g, 1e-3, 0.01, want_mask=True)
    ref = theta[None] - torch.clip(1e-3 * g, min=-0.01, max=0.01)
    assert torch.equal(out, ref)
    assert torch.equal(out_r, tf32_rn(ref))
    assert torch.equal(mask.bool(), (1e-3 * g).abs() <= 0.01)


def test_pos_embed_sine_matches_formula(ops):
    gen = torch.Generator(device="cuda").manual_seed(4)
    F_, h, w = 3, 19, 19
    mask = torch.zeros(F_, h, w, dtype=torch.bool, device="cuda")
    mask[1, :, 15:] = True
    mask[2, 17:, :] = True
    pos = ops.pos_embed_sine(mask)
    not_mask = ~mask
    y_embed = not_mask.cumsum(1, dtype=torch.float32)
    x_embed = not_mask.cumsum(2, dtype=torch.float32)
    eps, scale = 1e-6, 2 * 3.141592653589793
    y_embed = y_embed / (y_embed[:, -1:, :] + eps) * scale
    x_embed = x_embed / (x_embed[:, :, -1:] + eps) * scale
    dim_t = torch.arange(128, dtype=torch.float32, device="cuda")
    dim_t = 10000 ** (2 * (dim_t // 2) / 128)
    px = x_embed[:, :, :, None] / dim_t
    py = y_embed[:, :, :, None] / dim_t
    px = torch.stack((px[:, :, :, 0::2].sin(), px[:, :, :, 1::2].cos()), dim=4).flatten(3)
    py = torch.stack((py[:, :, :, 0::2].sin(), py[:, :, :, 1::2].cos()), dim=4).flatten(3)
    ref = torch.cat((py, px), dim=3).reshape(F_, h * w, 256)
    assert (pos - ref).abs().max().item() < 2e-5
    del gen


def test_matcher_cost_block_diagonal(ops):
    gen = torch.Generator(device="cuda").manual_seed(6)
    F_, Q, Cn = 5, 50, 1236
    logits = torch.randn(F_, Q, Cn, generator=gen, device="cuda")
    boxes = torch.rand(F_, Q, 4, generator=gen, device="cuda") * 0.5 + 0.1
    sizes = [3, 8, 5, 4, 6]
    T = sum(sizes)
    tb = torch.cat([torch.rand(T, 2, generator=gen, device="cuda") * 0.6 + 0.2,
                    torch.rand(T, 2, generator=gen, device="cuda") * 0.3 + 0.05], 1).contiguous()
    tl = torch.randint(1, 1235, (T,), generator=gen, device="cuda")
    off = torch.tensor([0] + list(torch.tensor(sizes).cumsum(0)), dtype=torch.int32, device="cuda")
    cost = ops.matcher_cost(logits, boxes, tb, tl, off, 1.0, 5.0, 2.0)

    def xyxy(b):
        cx, cy, w, h = b.unbind(-1)
        return torch.stack([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], -1)

    def giou(a, b):
        area1 = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
        area2 = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
        lt = torch.max(a[:, None, :2], b[:, :2])
        rb = torch.min(a[:, None, 2:], b[:, 2:])
        wh = (rb - lt).clamp(min=0)
        inter = wh[..., 0] * wh[..., 1]
        union = area1[:, None] + area2 - inter
        iou = inter / union
        lt = torch.min(a[:, None, :2], b[:, :2])
        rb = torch.max(a[:, None, 2:], b[:, 2:])
        wh = (rb - lt).clamp(min=0)
        area = wh[..., 0] * wh[..., 1]
        return iou - (area - union) / area

    prob = logits.flatten(0, 1).softmax(-1)
    ob = boxes.flatten(0, 1)
    Cfull = 5.0 * torch.cdist(ob, tb, p=1) - prob[:, tl] - 2.0 * giou(xyxy(ob), xyxy(tb))
    Cfull = Cfull.view(F_, Q, T)
    o = 0
    for f, n in enumerate(sizes):
        blk = cost[Q * o: Q * (o + n)].view(Q, n)
        assert (blk - Cfull[f, :, o:o + n]).abs().max().item() < 1e-5
        o += n


@pytest.mark.parametrize("geom", [(7, 2, 3, 1, 3, 30), (3, 1, 1, 1, 64, 19), (3, 2, 1, 1, 128, 38),
                                  (3, 1, 2, 2, 512, 19), (1, 2, 0, 1, 256, 38)])
def test_im2col_nhwc_matches_unfold(ops3, geom):
    """k, stride, pad, dilation, channels, size: the conv geometries of the ResNet-50-DC5 trunk."""
    k, stride, pad, dil, Cc, S = geom
    gen = torch.Generator(device="cuda").manual_seed(k * 100 + Cc)
    x = torch.randn(3, S, S, Cc, generator=gen, device="cuda")
    cols, Ho, Wo = ops3.im2col_nhwc(x, k, k, stride, pad, dil)
    ref = torch.nn.functional.unfold(x.permute(0, 3, 1, 2), (k, k), dilation=dil, padding=pad, stride=stride)
    ref = ref.view(3, Cc, k * k, Ho * Wo).permute(0, 3, 2, 1).reshape(3 * Ho * Wo, k * k * Cc)
    assert cols.shape[1] % 4 == 0 and torch.equal(cols[:, :k * k * Cc], ref)
    assert cols[:, k * k * Cc:].abs().sum() == 0


def test_maxpool_and_conv_as_gemm(ops3):
    gen = torch.Generator(device="cuda").manual_seed(21)
    x = torch.randn(2, 37, 37, 64, generator=gen, device="cuda")
    ref = torch.nn.functional.max_pool2d(x.permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1)
    assert torch.equal(ops3.maxpool3x3s2_nhwc(x), ref)
    # 3x3 dilated conv + bias + residual + ReLU-after-residual as one im2col + GEMM
    w = torch.randn(96, 64, 3, 3, generator=gen, device="cuda") * 0.05
    b = torch.randn(96, generator=gen, device="cuda")
    z = torch.randn(2, 37, 37, 96, generator=gen, device="cuda")
    cols, Ho, Wo = ops3.im2col_nhwc(x, 3, 3, 1, 2, 2)
    wm = w.permute(0, 2, 3, 1).reshape(96, -1).contiguous()
    y = ops3.matmul(cols, wm.t(), bias=b, act="relu", residual=z.view(-1, 96), act_after_residual=True)
    ref = torch.relu(torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), b.double(), 1, 2, 2)
                     .permute(0, 2, 3, 1) + z.double())
    assert rel(y.view(2, Ho, Wo, 96), ref) < X3_TOL


def test_gemm_trunk_matches_cudnn_fp32():
    """The im2col+GEMM trunk against cuDNN strict-fp32 convolutions on the same folded weights."""
    import interactron_b200 as ib
    from interactron_b200.backbone import run_backbone, run_backbone_gemm
    from interactron_b200.ops import CudaOps
    m = ib.build_model(ib.default_config("single_frame_baseline", weights="synthetic").MODEL).cuda().eval()
    body = m.model.backbone[0].body
    x = torch.randn(3, 3, 300, 300, generator=torch.Generator(device="cuda").manual_seed(2), device="cuda")
    a = run_backbone_gemm(body, x, CudaOps())
    b = run_backbone(body, x, tf32=False)
    assert a.shape == b.shape == (3, 19, 19, 2048)
    assert rel(a, b) < 2e-4       # 53 convolutions deep; cuDNN fp32 and tf32x3 each carry ~1e-5 per layer


@pytest.mark.parametrize("E,rows,n_out,k_in", [(1, 3610, 256, 256), (1, 3610, 2048, 256), (2, 3610, 256, 512),
                                               (1, 7220, 512, 512), (1, 4096, 256, 256)])
def test_gemm_split_k_weight_gradients(ops3, E, rows, n_out, k_in):
    """Weight-gradient GEMMs with few output tiles and a long K run split-K (K chunks as a batch dim +
    fixed-order colsum): same result as the one-chain launch and as fp64, into a strided slot of a flat
    gradient buffer, with and without accumulation; bit-identical across repeats."""
    gen = torch.Generator(device="cuda").manual_seed(rows + n_out)
    dy = torch.randn(E, rows, n_out, generator=gen, device="cuda")
    x = torch.randn(E, rows, k_in, generator=gen, device="cuda")
    want = dy.double().transpose(-1, -2) @ x.double()
    flat = torch.zeros(E, n_out * k_in + 8, device="cuda")
    out = flat[:, :n_out * k_in].view(E, n_out, k_in)
    n0 = ops3.n_split_k
    ops3.matmul(dy.transpose(-1, -2), x, out=out)
    assert ops3.n_split_k == n0 + 1, "split-K path not taken"
    assert rel(out, want) < X3_TOL and float(flat[:, n_out * k_in:].abs().sum()) == 0.0
    first = out.clone()
    ops3.matmul(dy.transpose(-1, -2), x, out=out)
    assert torch.equal(out, first)
    ops3.matmul(dy.transpose(-1, -2), x, out=out, accumulate=True)
    assert rel(out, 2 * want) < X3_TOL
    ops3.split_k = False
    try:
        one = torch.empty_like(out)
        ops3.matmul(dy.transpose(-1, -2), x, out=one)
        assert ops3.n_split_k == n0 + 3
    finally:
        ops3.split_k = True
    assert rel(first, one) < 1e-5
    if rows == 4096:                         # K-contiguous operands split too when the chunks stay 16-byte aligned
        a = torch.randn(1, n_out, rows, generator=gen, device="cuda")
        b = torch.randn(1, k_in, rows, generator=gen, device="cuda")
        o2 = torch.empty(1, n_out, k_in, device="cuda")
        ops3.matmul(a, b.transpose(-1, -2), out=o2)
        assert ops3.n_split_k == n0 + 4 and rel(o2, a.double() @ b.double().transpose(-1, -2)) < X3_TOL


def test_gemm_presplit_weights_bit_identical(ops3):
    """Static weights registered with their tf32 residuals (ops.register_presplit -> itn_gemm_desc_t::B_lo):
    the residual tile of B arrives by TMA instead of being recomputed in shared memory.  Same expression, so
    the results must be bit-identical to the plain launch - for whole weights, row slices, per-group stacks,
    transposed twins, and after an in-place update followed by a re-registration."""
    from interactron_b200.ops import CudaOps
    gen = torch.Generator(device="cuda").manual_seed(11)
    ops = CudaOps()                                   # own registry
    flat = torch.randn(1, 3 * 256 * 512 + 512 * 256, generator=gen, device="cuda")
    w = flat[:, :3 * 256 * 512].view(1, 768, 512)             # a packed in_proj-like weight [G=1, 3D, K]
    wt = flat[:, 3 * 256 * 512:].view(1, 512, 256)            # a transposed twin [G, K, N]
    x = torch.randn(1, 1805, 512, generator=gen, device="cuda")
    dy = torch.randn(1, 1805, 256, generator=gen, device="cuda")

    def run():
        return [ops.matmul(x, w.transpose(-1, -2)),                  # forward: x W^T, W [N, K] read K-major
                ops.matmul(x, w[:, 256:512].transpose(-1, -2)),      # a row slice (one of q / k / v)
                ops.matmul(dy, wt.transpose(-1, -2)),                # data gradient through the W^T twin, K-major
                ops.matmul(x, wt)]                                   # the twin read MN-major: plain path

    plain = [t.clone() for t in run()]
    assert ops.n_presplit == 0
    ops.register_presplit(flat)
    got = run()
    assert ops.n_presplit == 3          # the last product reads wt MN-major: no pre-split path for it
    for a, b in zip(got, plain):
        assert torch.equal(a, b)
    assert rel(got[0], x.double() @ w.double().transpose(-1, -2)) < X3_TOL
    # an activation operand outside every registered buffer is left alone
    n0 = ops.n_presplit
    ops.matmul(x, torch.randn(1, 512, 64, generator=gen, device="cuda"))
    assert ops.n_presplit == n0
    # in-place weight update: stale residuals until re-registered, exact again afterwards
    flat.mul_(1.37)
    ops.register_presplit(flat)
    after = ops.matmul(x, w.transpose(-1, -2))
    ops2 = CudaOps()
    assert torch.equal(after, ops2.matmul(x, w.transpose(-1, -2)))


def test_gemm_cta_pair_variant_is_bit_identical(ops3):
    """The cta_group::2 kernel (two CTAs share one 256-row MMA, each holding half of the B tile) issues the same
    products in the same order as the single-CTA kernel: results must be bit-identical.  Child processes, because
    the ITN_GEMM_PAIR switch is read once per process."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    child = (
        "import sys, torch\n"
        f"sys.path.insert(0, {root!r})\n"
        "from interactron_b200.ops import CudaOps\n"
        "ops = CudaOps(); torch.manual_seed(0)\n"
        "for M, N, K in ((40000, 256, 2048), (19000, 1236, 1024), (25000, 512, 1496)):\n"
        "    a = torch.randn(M, K, device='cuda'); w = torch.randn(N, K, device='cuda'); b = torch.randn(N, device='cuda')\n"
        "    r = torch.randn(M, N, device='cuda')\n"
        "    y = ops.matmul(a, w.t(), bias=b, residual=r)\n"
        "    ref = a.double() @ w.double().t() + b.double() + r.double()\n"
        "    print(repr(y.double().sum().item()), repr(y.abs().double().sum().item()), ((y.double() - ref).norm() / ref.norm()).item())\n")
    outs = []
    for pair in ("0", "1"):
        r = subprocess.run([sys.executable, "-c", child], env=dict(os.environ, ITN_GEMM_PAIR=pair), capture_output=True,
                           text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(r.stdout.strip().splitlines())
    assert outs[0] == outs[1] and len(outs[0]) == 3
    assert all(float(line.split()[2]) < 3e-6 for line in outs[0])


@pytest.mark.parametrize("a_mn", [False, True])
@pytest.mark.parametrize("b_mn", [False, True])
@pytest.mark.parametrize("shape", [(1805, 256, 256), (3000, 2048, 512), (1024, 512, 2048), (364, 32, 361), (250, 1236, 256),
                                   (777, 92, 4608)])
def test_gemm_tf32x3_split_accumulators_reach_fp32_error(a_mn, b_mn, shape):
    """ITN_PREC_TF32X3_SPLIT (CudaOps.split_acc): the residual products accumulate in tensor-memory columns of
    their own, so the main accumulator takes K/8 round-toward-zero accumulates instead of 3K/8.  Error against an
    fp64 product: no worse than twice torch's strict-fp32 GEMM on the same operands (or 4e-7 sqrt(K/256)), and below the default tf32x3
    mode; bias / relu / residual epilogues and batches agree with the default mode to the same accuracy."""
    from interactron_b200.ops import CudaOps
    o3, os_ = CudaOps(), CudaOps()
    os_.split_acc = True
    M, N, K = shape
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(K, M, generator=g, device="cuda").t() if a_mn else torch.randn(M, K, generator=g, device="cuda")
    b = torch.randn(N, K, generator=g, device="cuda").t() if not b_mn else torch.randn(K, N, generator=g, device="cuda")
    want = a.double() @ b.double()
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        e_torch = rel(a @ b, want)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    e_split, e_x3 = rel(os_.matmul(a, b), want), rel(o3.matmul(a, b), want)
    # (cuBLAS may split a long K over CTAs, which shortens its chains: allow the sqrt(K) growth of one chain)
    assert e_split < max(2.0 * e_torch, 4e-7 * (K / 256) ** 0.5), (e_split, e_torch)
    assert e_split <= e_x3 * 1.05, (e_split, e_x3)
    if N % 4 == 0:
        bias = torch.randn(N, generator=g, device="cuda")
        res = torch.randn(M, N, generator=g, device="cuda")
        y = os_.matmul(a, b, bias=bias, act="relu", residual=res, act_after_residual=True)
        assert rel(y, torch.relu(want + bias.double() + res.double())) < max(1e-6, 2.0 * e_split)
    a3 = torch.randn(3, 2, 200, 96, generator=g, device="cuda")
    b3 = torch.randn(3, 2, 96, 72, generator=g, device="cuda")
    assert rel(os_.matmul(a3, b3), a3.double() @ b3.double()) < 1e-6


@pytest.mark.parametrize("cols", [128, 256, 512])
@pytest.mark.parametrize("groups,rpg", [(1, 3000), (3, 250), (62, 50), (4, 1805), (2, 5)])
def test_layernorm_bwd_fused_one_launch(ops3, cols, groups, rpg):
    """itn_layernorm_bwd_fused: dx, dgamma, dbeta and colsum(dx) from one launch; same values as the
    dx / dgamma-dbeta / colsum kernels to fp32 rounding and as float64 autograd; strided outputs (slices of a
    flat gradient buffer); bit-identical across repeats (fixed summation order); optional outputs may be absent."""
    gen = torch.Generator(device="cuda").manual_seed(cols + groups + rpg)
    rows = groups * rpg
    x = torch.randn(rows, cols, generator=gen, device="cuda") * 2 + 0.5
    gamma = torch.randn(groups, cols, generator=gen, device="cuda")
    beta = torch.randn(groups, cols, generator=gen, device="cuda")
    dy = torch.randn(rows, cols, generator=gen, device="cuda")
    _, _, mean, rstd = ops3.layernorm_fwd(x, gamma, beta)
    flat = torch.zeros(groups, 3 * cols + 12, device="cuda")
    dg, db, ds = flat[:, :cols], flat[:, cols + 4:2 * cols + 4], flat[:, 2 * cols + 8:3 * cols + 8]
    assert ops3.fused_ln_bwd
    n0 = ops3.launch_count()
    dx, _ = ops3.layernorm_bwd(dy, x, mean, rstd, gamma, dgamma=dg, dbeta=db, dxsum=ds)
    assert ops3.launch_count() - n0 == 1
    xr = x.double().view(groups, -1, cols).requires_grad_(True)
    gr = gamma.double().requires_grad_(True)
    br = beta.double().requires_grad_(True)
    yr = torch.nn.functional.layer_norm(xr, (cols,)) * gr[:, None] + br[:, None]
    yr.backward(dy.double().view(groups, -1, cols))
    assert rel(dx, xr.grad.reshape(rows, cols)) < FP32_TOL
    assert rel(dg, gr.grad) < FP32_TOL and rel(db, br.grad) < FP32_TOL
    want_ds = xr.grad.sum(1)
    assert (ds.double() - want_ds).abs().max().item() < 1e-5 * xr.grad.abs().sum(1).max().item()
    pads = torch.cat([flat[:, cols:cols + 4], flat[:, 2 * cols + 4:2 * cols + 8], flat[:, 3 * cols + 8:]], 1)
    assert float(pads.abs().sum()) == 0.0
    first = flat.clone()
    dx2, _ = ops3.layernorm_bwd(dy, x, mean, rstd, gamma, dgamma=dg, dbeta=db, dxsum=ds)
    assert torch.equal(flat, first) and torch.equal(dx2, dx)
    # the three-kernel path computes the same thing
    ops3.fused_ln_bwd = False
    try:
        dg2, db2 = torch.empty(groups, cols, device="cuda"), torch.empty(groups, cols, device="cuda")
        dx3, _ = ops3.layernorm_bwd(dy, x, mean, rstd, gamma, dgamma=dg2, dbeta=db2)
    finally:
        ops3.fused_ln_bwd = True
    assert torch.equal(dx3, dx) and rel(dg2, dg) < 1e-6 and rel(db2, db) < 1e-6
    # only the bias gradient requested
    only = torch.zeros(groups, cols, device="cuda")
    dx4, _ = ops3.layernorm_bwd(dy, x, mean, rstd, gamma, dxsum=only)
    assert torch.equal(dx4, dx) and torch.equal(only, ds)


@pytest.mark.parametrize("case", ["full", "block", "grouped", "grouped_strided"])
def test_layernorm_fwd_plus_matches_layernorm_then_add(ops3, case):
    """itn_layernorm_fwd_plus: y and y + plus from one launch, bit-identical to layernorm_fwd followed by add,
    for every broadcast shape `add` is used with (full tensor, one block for all, one block per group)."""
    gen = torch.Generator(device="cuda").manual_seed(11)
    G, n, cols = 6, 50, 256
    x = torch.randn(G * n, cols, generator=gen, device="cuda") * 1.5 - 0.2
    gamma = torch.randn(G, cols, generator=gen, device="cuda")
    beta = torch.randn(G, cols, generator=gen, device="cuda")
    if case == "full":
        plus, shape = torch.randn(G * n, cols, generator=gen, device="cuda"), (G * n, cols)
    elif case == "block":
        plus, shape = torch.randn(1, n, cols, generator=gen, device="cuda"), (G, n, cols)
    elif case == "grouped":
        plus, shape = torch.randn(G, n, cols, generator=gen, device="cuda"), (G, n, cols)
    else:
        flat = torch.randn(G, n * cols + 1000, generator=gen, device="cuda")
        plus, shape = flat[:, 200:200 + n * cols].view(G, n, cols), (G, n, cols)
    assert ops3.fused_ln_plus
    n0 = ops3.launch_count()
    y, y_r, mean, rstd, yp = ops3.layernorm_fwd_plus(x, gamma, beta, plus, shape)
    assert ops3.launch_count() - n0 == 1
    y0, _, m0, r0 = ops3.layernorm_fwd(x, gamma, beta)
    yp0 = ops3.add(y0.view(shape), plus)
    assert torch.equal(y, y0) and torch.equal(mean, m0) and torch.equal(rstd, r0)
    assert torch.equal(yp.view(shape), yp0)
    want = y0.view(G, n, cols) + (plus.view(-1, n, cols) if case != "full" else plus.view(G, n, cols))
    assert torch.equal(yp.view(G, n, cols), want)


@pytest.mark.parametrize("geom", [(1, 8, 8, 32, 32, 3, 1, 1, 1), (2, 19, 19, 64, 64, 3, 1, 2, 2), (3, 38, 38, 64, 128, 3, 2, 1, 1),
                                  (3, 38, 38, 64, 128, 1, 2, 0, 1), (2, 75, 75, 64, 64, 3, 1, 1, 1),
                                  (1, 16, 16, 4, 64, 3, 1, 1, 1), (2, 30, 30, 4, 64, 7, 2, 3, 1), (3, 300, 300, 4, 64, 7, 2, 3, 1)])
def test_implicit_gemm_convolution_matches_explicit_im2col(ops3, geom):
    """Convolutions as ONE tcgen05 GEMM whose A tiles the TMA unit gathers in im2col mode (32-channel k-blocks, or -
    for the zero-padded RGB stem - 8 filter taps x 4 channels per k-block): bit-identical to the explicit
    im2col + GEMM path (same products, same order) and equal to float64 conv2d, on the trunk's geometries
    (3x3, dilated, strided 3x3 and 1x1, the 7x7/2 stem), with bias + ReLU (+ residual) fused."""
    import torch.nn.functional as F
    N, H, W, Cin, Cout, k, stride, pad, dil = geom
    g = torch.Generator(device="cuda").manual_seed(sum(geom))
    x = torch.randn(N, H, W, Cin, device="cuda", generator=g)
    w = torch.randn(Cout, k, k, Cin, device="cuda", generator=g) * (k * k * Cin) ** -0.5
    b = torch.randn(Cout, device="cuda", generator=g)
    wm = w.reshape(Cout, -1).contiguous()
    y, Ho, Wo = ops3.conv_gemm(x, wm, k, k, stride, pad, dil, bias=b, act="relu")
    a, Ho2, Wo2 = ops3.im2col_nhwc(x, k, k, stride, pad, dil)
    assert (Ho, Wo) == (Ho2, Wo2)
    y2 = ops3.matmul(a, wm.t(), bias=b, act="relu")
    assert torch.equal(y, y2)
    ref = F.conv2d(x.permute(0, 3, 1, 2).double(), w.permute(0, 3, 1, 2).double(), b.double(), stride, pad, dil)
    ref = ref.permute(0, 2, 3, 1).reshape(-1, Cout)
    assert rel(y, torch.relu(ref)) < 1e-5
    z = torch.randn(N * Ho * Wo, Cout, device="cuda", generator=g)
    y3, _, _ = ops3.conv_gemm(x, wm, k, k, stride, pad, dil, bias=b, act="relu", residual=z, act_after_residual=True)
    assert rel(y3, torch.relu(ref + z.double())) < 1e-5


def test_layernorm_bwd_workspace_growth_keeps_captured_graphs_valid():
    """A captured graph holds the address of the LayerNorm-backward workspace it was recorded with; when a later,
    larger problem makes the ops object allocate a bigger one, the old one must stay alive and usable."""
    from interactron_b200.ops import CudaOps
    o = CudaOps()
    g = torch.Generator(device="cuda").manual_seed(3)
    rows, cols, groups = 6 * 250, 256, 6
    x = torch.randn(rows, cols, generator=g, device="cuda")
    dy = torch.randn(rows, cols, generator=g, device="cuda")
    gamma = torch.randn(groups, cols, generator=g, device="cuda")
    beta = torch.zeros(groups, cols, device="cuda")
    _, _, mean, rstd = o.layernorm_fwd(x, gamma, beta)
    dg, db, ds = (torch.zeros(groups, cols, device="cuda") for _ in range(3))
    want_dx, _ = o.layernorm_bwd(dy, x, mean, rstd, gamma, dgamma=dg, dbeta=db, dxsum=ds)       # eager: sizes the workspace
    want = [t.clone() for t in (want_dx, dg, db, ds)]
    small_ws = o._ln_ws
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        got_dx, _ = o.layernorm_bwd(dy, x, mean, rstd, gamma, dgamma=dg, dbeta=db, dxsum=ds)
    # a problem whose workspace does not fit the first allocation (2048 groups x 3 x 512 floats)
    G2 = 2048
    x2 = torch.randn(G2 * 2, 512, generator=g, device="cuda")
    ga2 = torch.ones(G2, 512, device="cuda")
    _, _, m2, r2 = o.layernorm_fwd(x2, ga2, torch.zeros(G2, 512, device="cuda"))
    o.layernorm_bwd(torch.randn(G2 * 2, 512, generator=g, device="cuda"), x2, m2, r2, ga2,
                    dgamma=torch.zeros(G2, 512, device="cuda"), dbeta=torch.zeros(G2, 512, device="cuda"))
    assert o._ln_ws is not small_ws and any(w is small_ws for w in o._ln_ws_keep)
    junk = [torch.full((1 << 20,), float("nan"), device="cuda") for _ in range(16)]      # would land in a freed workspace
    for t in (dg, db, ds):
        t.zero_()
    graph.replay()
    torch.cuda.synchronize()
    for a, b in zip((got_dx, dg, db, ds), want):
        assert torch.equal(a, b)
    del junk
