"""Deterministic synthetic weights and episodes (no dataset / checkpoint is available offline).

Both the new path and the oracle (the reference run on CPU) load the *same*
state_dict produced here, and consume the same seeded episodes.
Everything is generated on the CPU generator so the values are identical on
every box with the same torch build.
"""
import math
import zlib

import torch

WEIGHT_SEED = 1234
ATTN_SHARPEN = 3.0   # q/k/v projection gain: makes the random-init attention maps non-uniform
BN3_DAMP = 0.5   # damping of the last BN of each bottleneck: keeps random-init trunk features O(1)


def _fans(shape):
    if len(shape) < 2:
        return shape[0], shape[0]
    rf = 1
    for s in shape[2:]:
        rf *= s
    return shape[1] * rf, shape[0] * rf


def _init_tensor(key, t, gen):
    """Random init with reference-like scales.  Unlike a fresh reference model, no tensor is
    left at exactly 0/1 (biases, LayerNorm affines, position tables), so that parity tests are
    sensitive to every term of the computation."""
    shape = tuple(t.shape)
    u = lambda a: (torch.rand(shape, generator=gen, dtype=torch.float32) * 2 - 1) * a
    n = lambda s: torch.randn(shape, generator=gen, dtype=torch.float32) * s
    leaf = key.rsplit(".", 1)[-1]
    if key.endswith("attn.mask"):
        return torch.ones(shape)
    if leaf == "pos_embed":  # fusion B fixed sincos table: keep the deterministic value
        return t.clone()
    if "backbone" in key:
        if leaf == "running_var":
            return 1.0 + torch.rand(shape, generator=gen) * 0.2
        if leaf == "running_mean":
            return u(0.05)
        if "bn" in key or "downsample.1" in key:
            if leaf == "weight":
                # the last BN of each bottleneck is damped so the random-init trunk keeps O(1) features
                return (BN3_DAMP if ".bn3." in key else 1.0) * (1.0 + u(0.1))
            if leaf == "bias":
                return u(0.05)
        fan_in, fan_out = _fans(shape)
        return n(math.sqrt(2.0 / fan_out))          # kaiming-normal, fan_out (torchvision resnet)
    if leaf in ("seq_pos_embed", "pos_emb", "query_embed") and "detector" not in key and t.dim() == 3:
        return n(0.02)
    if key.endswith("query_embed.weight"):
        return n(1.0)
    if leaf == "action_tokens":
        return u(1.0 / math.sqrt(shape[-1]))
    is_norm = ".norm" in key or ".ln" in key or key.startswith("norm") or "ln_f" in key
    if is_norm:
        return 1.0 + u(0.1) if leaf == "weight" else u(0.05)
    if t.dim() >= 2:
        if ".model." in "." + key and leaf == "weight":     # GPT linears: N(0, 0.02)
            sharp = ATTN_SHARPEN if (".attn.key." in key or ".attn.query." in key) else 1.0
            return n(0.02 * sharp)
        fan_in, fan_out = _fans(shape)
        sharp = ATTN_SHARPEN if leaf == "in_proj_weight" else 1.0
        return u(sharp * math.sqrt(6.0 / (fan_in + fan_out)))   # xavier-uniform
    return u(0.05)                                            # biases


def synthetic_state_dict(model, seed=WEIGHT_SEED):
    """state_dict with every tensor drawn from a CPU generator seeded by (seed, key name), so a
    tensor gets the same values whichever model class holds it (`detector.*` of the adaptive
    models == `model.*` of the plain DETR baseline).  `criterion.*` buffers are left untouched."""
    out = {}
    for key, t in model.state_dict().items():
        if key.startswith("criterion.") or not t.is_floating_point():
            out[key] = t.clone()
            continue
        canon = "detector." + key[len("model."):] if key.startswith("model.") else key
        gen = torch.Generator(device="cpu").manual_seed(seed * 1_000_003 + zlib.crc32(canon.encode()))
        out[key] = _init_tensor(canon, t, gen).to(t.dtype).reshape(t.shape)
    return out


def synthetic_episode(episode_id, frames=5, res=300, with_targets=True):
    """One reference-format `data` dict with batch size 1 (reference utils/storage_utils.py:53-64):
    frames ~ N(0,1) [1,5,3,res,res]; masks int64 zeros; 3..8 targets per frame with labels in
    1..1234 and valid cxcywh boxes (centre U(.2,.8), size U(.05,.35)); actions U{0..3}."""
    gen = torch.Generator(device="cpu").manual_seed(10_000 + episode_id)
    data = {
        "frames": torch.randn(1, frames, 3, res, res, generator=gen),
        "masks": torch.zeros(1, frames, res, res, dtype=torch.long),
        "actions": torch.randint(0, 4, (1, frames), generator=gen),
        "episode_ids": torch.tensor([episode_id]),
        "initial_image_path": [f"ep{episode_id}"],
    }
    if with_targets:
        cats, boxes = [], []
        for _ in range(frames):
            nt = int(torch.randint(3, 9, (1,), generator=gen))
            cats.append(torch.randint(1, 1235, (nt,), generator=gen))
            centre = torch.rand(nt, 2, generator=gen) * 0.6 + 0.2
            size = torch.rand(nt, 2, generator=gen) * 0.3 + 0.05
            boxes.append(torch.cat([centre, size], dim=1))
        data["category_ids"] = [cats]
        data["boxes"] = [boxes]
        data["object_ids"] = [[torch.arange(len(c)) for c in cats]]
    return data


def masked_episode(episode_id, frames=5, res=300, with_targets=True):
    """synthetic_episode with padded frames: frame 1 padded on the right quarter, frame 3 on the right
    quarter and the bottom sixth (mask = 1, pixels zeroed, as DETR's NestedTensor padding does).  The
    dataset never produces such masks (datasets/sequence_dataset.py:56) but the interface carries them
    into the key-padding masks and the sine position embedding."""
    data = synthetic_episode(episode_id, frames, res, with_targets)
    m = data["masks"]
    m[0, 1, :, res - res // 4:] = 1
    m[0, 3, :, res - res // 4:] = 1
    m[0, 3, res - res // 6:, :] = 1
    data["frames"] = data["frames"] * (m == 0)[:, :, None].to(data["frames"].dtype)
    return data


def collate_episodes(episodes):
    """Stack batch-1 episodes into one batch-B `data` dict."""
    out = {
        "frames": torch.cat([e["frames"] for e in episodes], 0),
        "masks": torch.cat([e["masks"] for e in episodes], 0),
        "actions": torch.cat([e["actions"] for e in episodes], 0),
        "episode_ids": torch.cat([e["episode_ids"] for e in episodes], 0),
        "initial_image_path": [p for e in episodes for p in e["initial_image_path"]],
    }
    for k in ("category_ids", "boxes", "object_ids"):
        if k in episodes[0]:
            out[k] = [x for e in episodes for x in e[k]]
    return out
