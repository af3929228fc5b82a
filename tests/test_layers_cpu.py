"""Host logic of the shared layer blocks on the float64 simulator backend (no GPU)."""
import torch

from interactron_b200 import layers
from oracle.sim_ops import SimOps


def test_l2_chunked_attention_is_identical():
    """Chunking the batch so the score tensors stay L2-resident must not change a single bit."""
    torch.manual_seed(0)
    B, Lq, Lk, nh, hd = 7, 9, 13, 2, 4
    D = nh * hd
    ops = SimOps(torch.float64)
    q, k, v = (torch.randn(B, L, D, dtype=torch.float64) for L in (Lq, Lk, Lk))
    dO = torch.randn(B, Lq, D, dtype=torch.float64)
    kmask = (torch.rand(B, Lk) < 0.2).to(torch.uint8)

    def run():
        o, P = layers.attention_fwd(ops, q, k, v, B, Lq, Lk, nh, hd, 0.5, kmask)
        dq, dk, dv = torch.zeros_like(q), torch.zeros_like(k), torch.zeros_like(v)
        layers.attention_bwd(ops, dO, q, k, v, P, B, Lq, Lk, nh, hd, 0.5, dq, dk, dv)
        return o, P, dq, dk, dv

    ref = run()
    per_batch = nh * Lq * layers.pad4(Lk) * 8
    assert layers._l2_chunks(ops, B, per_batch, 1) == [(0, B)]
    ops.attn_l2_mb = 2.5 * per_batch / (1 << 20)            # 2 batches per forward chunk, 1 per backward chunk
    assert len(layers._l2_chunks(ops, B, per_batch * 1, 1)) > 1
    for a, b in zip(ref, run()):
        assert torch.equal(a, b)
    # autograd cross-check of the (chunked) backward
    qa, ka, va = (t.clone().requires_grad_() for t in (q, k, v))
    s = torch.einsum("bqhd,bkhd->bhqk", qa.view(B, Lq, nh, hd), ka.view(B, Lk, nh, hd)) * 0.5
    s = s.masked_fill(kmask.bool()[:, None, None, :], float("-inf"))
    o = torch.einsum("bhqk,bkhd->bqhd", s.softmax(-1), va.view(B, Lk, nh, hd)).reshape(B, Lq, D)
    g = torch.autograd.grad((o * dO).sum(), (qa, ka, va))
    assert (o - ref[0]).abs().max() < 1e-12
    for a, b in zip(g, ref[2:]):
        assert (a - b).abs().max() < 1e-12


def test_host_mask_sampling_is_interpolate_nearest():
    """episode.sample_masks_host == the reference's F.interpolate(mask, size=feature map) (nearest,
    detr_models/backbone.py:77) on the host, and trunk_hw == the real ResNet-50-DC5 output size."""
    import torch.nn.functional as F
    import torchvision
    from interactron_b200.episode import sample_masks_host, trunk_hw
    torch.manual_seed(0)
    for (H, W) in ((300, 300), (224, 320), (301, 299), (512, 640)):
        h, w = trunk_hw(H, W)
        m = (torch.rand(2, 5, H, W) < 0.3).long() * 7
        ref = F.interpolate(m.flatten(0, 1)[None].float(), size=(h, w)).to(torch.bool)[0].reshape(2, 5, h, w)
        got = sample_masks_host(m)
        assert got.dtype == torch.uint8 and torch.equal(got.bool(), ref), (H, W)
    body = torchvision.models.resnet50(replace_stride_with_dilation=[False, False, True])
    body = torch.nn.Sequential(*list(body.children())[:-2]).eval()
    for (H, W) in ((300, 300), (224, 320), (301, 299)):
        with torch.no_grad():
            o = body(torch.zeros(1, 3, H, W))
        assert tuple(o.shape[-2:]) == trunk_hw(H, W), (H, W)


def test_split_k_decomposition_host_logic():
    """ops.CudaOps split-K: the K range as one more batch dim of strided views + fixed-order sum of the
    partials == the plain product (host-side view arithmetic; the kernels are replaced by torch here)."""
    from interactron_b200.ops import CudaOps, _as4d

    class Fake:
        n_split_k = 0

        def empty(self, *shape):
            return torch.full(shape, float("nan"), dtype=torch.float64)

        def matmul(self, a, b, out=None, _nosplit=False):
            assert _nosplit
            out.copy_(torch.matmul(a, b))

        colsum = SimOps.colsum
        calls = 0

    torch.manual_seed(1)
    fake = Fake()
    for E, rows, n_out, k_in in ((1, 3610, 256, 256), (2, 3610, 256, 512), (1, 1805 * 3, 64, 96)):
        dy = torch.randn(E, rows, n_out, dtype=torch.float64)
        x = torch.randn(E, rows, k_in, dtype=torch.float64)
        a4, b4 = _as4d(dy.transpose(-1, -2)), _as4d(x)                 # dW = dy^T x, both operands MN-major
        flat = torch.zeros(E, n_out * k_in + 8, dtype=torch.float64)
        out = flat[:, :n_out * k_in].view(E, n_out, k_in)              # a slot of a flat gradient buffer
        S = CudaOps._split_k_factor(fake, a4, b4, out, n_out, k_in, rows, 1, E)
        assert S >= 4 and rows % S == 0, (S, rows)
        CudaOps._matmul_split_k(fake, a4, b4, out, S, False)
        want = dy.transpose(-1, -2) @ x
        assert (out - want).abs().max() < 1e-10 and float(flat[:, n_out * k_in:].abs().sum()) == 0.0
        CudaOps._matmul_split_k(fake, a4, b4, out, S, True)             # accumulate: out += product
        assert (out - 2 * want).abs().max() < 1e-10
    # K-contiguous operands need 16-byte aligned chunk starts; big problems and small K are left alone
    a4, b4 = _as4d(torch.zeros(1, 256, 3610)), _as4d(torch.zeros(1, 3610, 256))
    assert CudaOps._split_k_factor(fake, a4, b4, torch.zeros(1, 256, 256), 256, 256, 3610, 1, 1) == 1
    a4 = _as4d(torch.zeros(1, 256, 4096))
    assert CudaOps._split_k_factor(fake, a4, _as4d(torch.zeros(1, 4096, 256)), torch.zeros(1, 256, 256), 256, 256, 4096, 1, 1) >= 4
    big = _as4d(torch.zeros(32, 2048, 1805).transpose(-1, -2).transpose(-1, -2))
    assert CudaOps._split_k_factor(fake, big, _as4d(torch.zeros(32, 1805, 256)), torch.zeros(32, 2048, 256), 2048, 256, 1805, 1, 32) == 1
