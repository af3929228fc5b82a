"""Parameter containers whose `state_dict()` layout equals the reference's.

These modules only *hold* weights (same key names, shapes and registration order
as allenai/interactron, so reference checkpoints load unchanged):
  detector.*  — DETR-DC5/R50        (reference models/detr_models/detr.py:22-46,314-341)
  fusion.*    — fusion A (GPT)      (reference models/transformer.py:33-45, models/gpt.py:82-103)
              — fusion B (decoder)  (reference models/new_transformer.py:12-32)
Their `forward` is never used for arithmetic: the hot path reads the tensors and
runs the sm_100a kernels (see detr_t.py / fusion_a.py / fusion_b.py).  Only the
frozen ResNet backbone is executed through PyTorch/cuDNN (backbone.py).
"""
import math

import torch
from torch import nn


class FrozenBN(nn.Module):
    """Affine + statistics kept as buffers (never trained), as in DETR's frozen BatchNorm."""

    def __init__(self, n):
        super().__init__()
        self.register_buffer("weight", torch.ones(n))
        self.register_buffer("bias", torch.zeros(n))
        self.register_buffer("running_mean", torch.zeros(n))
        self.register_buffer("running_var", torch.ones(n))

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        state_dict.pop(prefix + "num_batches_tracked", None)
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)

    def scale_shift(self, eps=1e-5):
        scale = self.weight * (self.running_var + eps).rsqrt()
        return scale, self.bias - self.running_mean * scale

    def forward(self, x):
        scale, shift = self.scale_shift()
        return x * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)


class _Body(nn.Module):
    """Holds the ResNet-50 (dilated last stage) trunk under the key `body`."""

    def __init__(self):
        super().__init__()
        import torchvision
        from torchvision.models._utils import IntermediateLayerGetter
        net = torchvision.models.resnet50(weights=None, replace_stride_with_dilation=[False, False, True],
                                          norm_layer=FrozenBN)
        self.body = IntermediateLayerGetter(net, return_layers={"layer4": "0"})
        self.num_channels = 2048


class _NoParams(nn.Module):
    """Placeholder for the parameter-free sine position embedding (backbone.1)."""


class MLPHolder(nn.Module):
    def __init__(self, d_in, d_hidden, d_out, n_layers):
        super().__init__()
        self.num_layers = n_layers
        dims = [d_in] + [d_hidden] * (n_layers - 1) + [d_out]
        self.layers = nn.ModuleList(nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:]))


class EncoderLayerHolder(nn.Module):
    def __init__(self, d, heads, ffn):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(d, heads, dropout=0.1)
        self.linear1 = nn.Linear(d, ffn)
        self.dropout = nn.Dropout(0.1)
        self.linear2 = nn.Linear(ffn, d)
        self.norm1 = nn.LayerNorm(d)
        self.norm2 = nn.LayerNorm(d)
        self.dropout1 = nn.Dropout(0.1)
        self.dropout2 = nn.Dropout(0.1)


class DecoderLayerHolder(nn.Module):
    def __init__(self, d, heads, ffn):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(d, heads, dropout=0.1)
        self.multihead_attn = nn.MultiheadAttention(d, heads, dropout=0.1)
        self.linear1 = nn.Linear(d, ffn)
        self.dropout = nn.Dropout(0.1)
        self.linear2 = nn.Linear(ffn, d)
        self.norm1 = nn.LayerNorm(d)
        self.norm2 = nn.LayerNorm(d)
        self.norm3 = nn.LayerNorm(d)
        self.dropout1 = nn.Dropout(0.1)
        self.dropout2 = nn.Dropout(0.1)
        self.dropout3 = nn.Dropout(0.1)


class _Stack(nn.Module):
    def __init__(self, make_layer, n, norm=None):
        super().__init__()
        self.layers = nn.ModuleList(make_layer() for _ in range(n))
        self.norm = norm


class _DetrTransformer(nn.Module):
    def __init__(self, d=256, heads=8, ffn=2048, n_enc=6, n_dec=6):
        super().__init__()
        self.encoder = _Stack(lambda: EncoderLayerHolder(d, heads, ffn), n_enc, None)
        self.decoder = _Stack(lambda: DecoderLayerHolder(d, heads, ffn), n_dec, nn.LayerNorm(d))
        self.d_model, self.nhead = d, heads


class DetectorHolder(nn.Module):
    """DETR weights: 6+6 post-norm transformer (d=256, 8 heads, ffn 2048), 50 queries."""

    def __init__(self, num_classes):
        super().__init__()
        self.num_queries = 50
        self.transformer = _DetrTransformer()
        d = self.transformer.d_model
        self.class_embed = nn.Linear(d, num_classes + 1)
        self.bbox_embed = MLPHolder(d, d, 4, 3)
        self.query_embed = nn.Embedding(self.num_queries, d)
        self.input_proj = nn.Conv2d(2048, d, kernel_size=1)
        self.backbone = nn.Sequential(_Body(), _NoParams())
        # Decision D1: the backbone is frozen, so it is not a fast weight.
        self.backbone.requires_grad_(False)


class _GPTAttnHolder(nn.Module):
    def __init__(self, d, block_size):
        super().__init__()
        self.key = nn.Linear(d, d)
        self.query = nn.Linear(d, d)
        self.value = nn.Linear(d, d)
        self.attn_drop = nn.Dropout(0.1)
        self.resid_drop = nn.Dropout(0.1)
        self.proj = nn.Linear(d, d)
        # the reference checkpoints carry this all-ones [1,1,T,T] buffer; it is accepted and emitted
        # but never read (an all-ones mask is a no-op, reference models/gpt.py:35-36,49)
        self.register_buffer("mask", torch.ones(1, 1, block_size, block_size))


class _GPTBlockHolder(nn.Module):
    def __init__(self, d, block_size):
        super().__init__()
        self.ln1 = nn.LayerNorm(d)
        self.ln2 = nn.LayerNorm(d)
        self.attn = _GPTAttnHolder(d, block_size)
        self.mlp = nn.Sequential(nn.Linear(d, 4 * d), nn.GELU(), nn.Linear(4 * d, d), nn.Dropout(0.1))


class _GPTHolder(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        d = cfg.EMBEDDING_DIM
        self.pos_emb = nn.Parameter(torch.zeros(1, 255, d))
        self.seq_pos_embed = nn.Parameter(torch.zeros(1, 2060, d))
        self.drop = nn.Dropout(cfg.EMBEDDING_PDROP)
        self.blocks = nn.Sequential(*[_GPTBlockHolder(d, cfg.BLOCK_SIZE) for _ in range(cfg.NUM_LAYERS)])
        self.ln_f = nn.LayerNorm(d)
        self.head = nn.Linear(d, cfg.OUTPUT_SIZE, bias=False)
        self.block_size = cfg.BLOCK_SIZE


class FusionAHolder(nn.Module):
    """Fusion A weights: token embeddings + 4-layer pre-LN GPT (d=512) + decoders."""

    def __init__(self, cfg):
        super().__init__()
        d = cfg.EMBEDDING_DIM
        self.img_feature_embedding = nn.Linear(cfg.IMG_FEATURE_SIZE, d)
        self.prediction_embedding = nn.Linear(cfg.BOX_EMB_SIZE + cfg.NUM_CLASSES + 5, d)
        self.model = _GPTHolder(cfg)
        self.box_decoder = MLPHolder(cfg.OUTPUT_SIZE, 256, 4, 3)
        self.logit_decoder = nn.Linear(cfg.OUTPUT_SIZE, cfg.NUM_CLASSES + 1)
        self.loss_decoder = MLPHolder(cfg.OUTPUT_SIZE, 512, 1, 3)
        self.action_decoder = MLPHolder(cfg.OUTPUT_SIZE, 512, 4, 3)
        self.action_tokens = nn.Parameter(torch.empty(1, 5, d))
        nn.init.kaiming_uniform_(self.action_tokens, a=math.sqrt(5))


class FusionBHolder(nn.Module):
    """Fusion B weights: token embeddings + 4 DETR decoder layers (d=512) + decoders."""

    def __init__(self, cfg):
        super().__init__()
        d = cfg.EMBEDDING_DIM
        self.img_feature_embedding = nn.Linear(cfg.IMG_FEATURE_SIZE, d)
        self.prediction_embedding = nn.Linear(cfg.BOX_EMB_SIZE + cfg.NUM_CLASSES + 5, d)
        self.box_decoder = MLPHolder(cfg.OUTPUT_SIZE, 512, 4, 3)
        self.logit_decoder = nn.Linear(cfg.OUTPUT_SIZE, cfg.NUM_CLASSES + 1)
        self.loss_decoder = MLPHolder(cfg.OUTPUT_SIZE, 512, 1, 3)
        self.action_decoder = MLPHolder(cfg.OUTPUT_SIZE, 512, 4, 3)
        self.action_tokens = nn.Parameter(torch.empty(1, 5, d))
        nn.init.kaiming_uniform_(self.action_tokens, a=math.sqrt(5))
        self.transformer = _Stack(lambda: DecoderLayerHolder(d, cfg.NUM_HEADS, 2048), cfg.NUM_LAYERS,
                                  nn.LayerNorm(d))
        self.embed_dim = d
        self.pos_embed = nn.Parameter(sincos_memory_pos(d), requires_grad=False)
        self.query_embed = nn.Parameter(torch.zeros(1, 255, d))


def _sincos_1d(dim, positions):
    """[len(positions), dim]: sin block then cos block, frequencies 1/10000^(i/(dim/2))."""
    omega = 1.0 / (10000.0 ** (torch.arange(dim // 2, dtype=torch.float64) / (dim / 2.0)))
    ang = positions.double().reshape(-1)[:, None] * omega[None, :]
    return torch.cat([ang.sin(), ang.cos()], dim=1)


def sincos_memory_pos(d, grid=19, frames=5):
    """Fixed position code of fusion B's memory tokens (reference models/new_transformer.py:60-73):
    first d/2 channels = 2-D sincos of the 19x19 cell, last d/2 = 1-D sincos of the frame index."""
    half = d // 2
    ys, xs = torch.meshgrid(torch.arange(grid), torch.arange(grid), indexing="ij")
    # the 2-D code spends half//2 channels on the column index, then half//2 on the row index
    emb_2d = torch.cat([_sincos_1d(half // 2, xs), _sincos_1d(half // 2, ys)], dim=1)
    emb_f = _sincos_1d(half, torch.arange(frames))
    pos = torch.zeros(1, frames * grid * grid, d, dtype=torch.float64)
    for i in range(frames):
        blk = pos[0, i * grid * grid:(i + 1) * grid * grid]
        blk[:, :half] = emb_2d
        blk[:, half:] = emb_f[i]
    return pos.float()


def fast_weight_items(detector):
    """(name, parameter) pairs of the detector's fast weights theta, in the reference's order.

    The reference collects, depth-first in registration order, the `requires_grad`
    parameters owned *directly* by modules that have no sub-modules
    (utils/meta_utils.py:5-24).  A module with children contributes nothing of its
    own, which is why nn.MultiheadAttention.in_proj_* (it owns `out_proj`) is never
    adapted.  With the backbone frozen (D1) this yields 157 tensors / 14,798,296 elements.
    """
    out = []

    def walk(mod, prefix):
        kids = list(mod.named_children())
        if not kids:
            for pname, p in mod._parameters.items():
                if p is not None and p.requires_grad:
                    out.append((prefix + pname, p))
            return
        for cname, child in kids:
            walk(child, prefix + cname + ".")

    walk(detector, "")
    return out
