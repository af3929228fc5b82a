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
