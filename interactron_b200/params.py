"""Flat parameter buffers.

The reference clones / detaches / re-installs ~200 separate tensors around every
episode (utils/meta_utils.py:48-111).  Here the fast weights theta live in ONE flat
fp32 buffer (plus a TF32-rounded twin used as GEMM weights), so clone/detach/set
become pointer arithmetic, the SGD step is one kernel over the buffer, and
per-episode fast weights are simply rows of a [episodes, n] matrix.
"""
import torch


class ParamPack:
    """Names, shapes and offsets of a list of tensors inside a flat buffer."""

    def __init__(self, items):
        self.names, self.shapes, self.offsets, self.sizes, self.pads = [], {}, {}, {}, {}
        off = 0
        for name, t in items:
            n = t.numel()
            self.names.append(name)
            self.shapes[name] = tuple(t.shape)
            self.offsets[name] = off
            self.sizes[name] = n
            self.pads[name] = (-n) % 4          # every tensor starts 16-byte aligned (TMA / float4)
            off += n + self.pads[name]
        self.numel = off

    def __contains__(self, name):
        return name in self.offsets

    def pack(self, tensors, device=None, dtype=torch.float32):
        """Concatenate the (detached) tensors into a new flat [numel] buffer."""
        parts = []
        for name, t in zip(self.names, tensors):
            parts.append(t.detach().reshape(-1).to(dtype))
            if self.pads[name]:
                parts.append(torch.zeros(self.pads[name], dtype=dtype, device=t.device))
        flat = torch.cat(parts)
        return flat.to(device) if device is not None else flat

    def view(self, flat, name):
        """flat [..., numel] -> view [..., *shape] of tensor `name`."""
        off, n = self.offsets[name], self.sizes[name]
        return flat[..., off:off + n].reshape(*flat.shape[:-1], *self.shapes[name])

    def unpack(self, flat):
        return {n: self.view(flat, n) for n in self.names}


class Weights:
    """Named weight views over one or more (pack, flat, flat_rounded) triples.

    `p(name)` is the full-precision tensor (biases, LayerNorm affines, embeddings);
    `w(name)` the TF32-rounded twin that the tensor cores read.  Every view has a
    leading group dim G (1 = shared by all episodes, E = one fast-weight set per episode).
    """

    def __init__(self, *triples):
        self.triples = triples

    def _find(self, name):
        for pack, flat, flat_r in self.triples:
            if name in pack:
                return pack, flat, flat_r
        raise KeyError(name)

    def p(self, name):
        pack, flat, _ = self._find(name)
        return pack.view(flat, name)

    def w(self, name):
        pack, _, flat_r = self._find(name)
        return pack.view(flat_r, name)


def detector_packs(detector):
    """(theta_pack, theta_params, psi_pack, psi_params) of a DetectorHolder.

    theta = the fast weights in reference order (modules.fast_weight_items); psi = the remaining
    transformer parameters (all `in_proj_*`), used by both passes but never adapted."""
    from .modules import fast_weight_items
    theta = fast_weight_items(detector)
    theta_ids = {id(p) for _, p in theta}
    psi = [(n, p) for n, p in detector.named_parameters()
           if id(p) not in theta_ids and not n.startswith("backbone.")]
    return (ParamPack(theta), [p for _, p in theta], ParamPack(psi), [p for _, p in psi])


def flat_weights(ops, pack, params):
    """-> (flat [1,n] fp32, TF32-rounded twin) on the ops device."""
    flat = pack.pack(params, device=ops.device).unsqueeze(0)
    return flat, ops.round_tf32(flat)
