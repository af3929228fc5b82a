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

    def matrix_shape(self, name):
        """(N_out, K_in) of a >= 2-D weight (trailing dims flattened, e.g. the 1x1 conv)."""
        shp = self.shapes[name]
        k = 1
        for d in shp[1:]:
            k *= d
        return shp[0], k

    def view_t(self, flat_t, name):
        """flat_t [..., numel] -> W^T view [..., K_in, N_out] of the transposed twin buffer."""
        off, n = self.offsets[name], self.sizes[name]
        N, K = self.matrix_shape(name)
        return flat_t[..., off:off + n].reshape(*flat_t.shape[:-1], K, N)

    def transpose_into(self, ops, flat, flat_t):
        """Fill flat_t with the per-tensor transposes of every >= 2-D tensor of flat ([G, numel])."""
        for name in self.names:
            if len(self.shapes[name]) < 2:
                continue
            N, K = self.matrix_shape(name)
            src = self.view(flat, name).reshape(flat.shape[0], N, K)
            ops.transpose_(self.view_t(flat_t, name), src)
        return flat_t


class Weights:
    """Named weight views over one or more (pack, flat, flat_rounded[, flat_transposed]) tuples.

    `p(name)`   full-precision tensor (biases, LayerNorm affines, embeddings);
    `w(name)`   GEMM weight [G, N_out, K_in...] (TF32-rounded twin in single-pass mode);
    `bwd(name)` the b-operand [G, N_out, K_in] of a data-gradient GEMM `dy @ W`, served from the
                transposed twin W^T when there is one so the tensor core reads it K-major.
    Every view has a leading group dim G (1 = shared by all episodes, E = per-episode theta')."""

    def __init__(self, *tuples):
        self.tuples = [t if len(t) == 4 else (*t, None) for t in tuples]

    def _find(self, name):
        for t in self.tuples:
            if name in t[0]:
                return t
        raise KeyError(name)

    def p(self, name):
        pack, flat, _, _ = self._find(name)
        return pack.view(flat, name)

    def w(self, name):
        pack, _, flat_r, _ = self._find(name)
        return pack.view(flat_r, name)

    def bwd(self, name, lo=None, hi=None):
        pack, _, flat_r, flat_t = self._find(name)
        if flat_t is not None:
            wt = pack.view_t(flat_t, name)
            if lo is not None:
                wt = wt[..., lo:hi]
            return wt.transpose(-1, -2)
        N, K = pack.matrix_shape(name)
        w = pack.view(flat_r, name).reshape(flat_r.shape[0], N, K)
        return w[..., lo:hi, :] if lo is not None else w


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
