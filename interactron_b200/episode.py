"""The inner loop: adapt on a 5-frame episode, then re-detect (reference
models/interactron.py:31-59, models/interactron_random.py:27-55, utils/meta_utils.py).

For E episodes at once (the reference handles one):
  1. frozen backbone (PyTorch/cuDNN) on the E*S frames, once per episode (decision D1);
  2. DETR transformer forward with the shared weights theta              (detr_t)
  3. fusion forward -> learned loss = ||loss_decoder(...)||_2 per episode (fusion)
  4. backward: d learned_loss / d theta for every episode                 (fusion + detr_t)
  5. theta'_e = theta - clip(lr * g_e, +-0.01): ONE fused kernel over the flat buffer
  6. DETR transformer forward with the per-episode fast weights theta'_e (grouped GEMMs).
theta lives in one flat buffer, so the reference's clone/detach/set_parameters bookkeeping
(utils/meta_utils.py:48-111) is pointer arithmetic here.
"""
import os

import torch

from . import detr_t, fusion

from .backbone import run_backbone, run_backbone_gemm
from .layers import GradSink
from .params import ParamPack, Weights, detector_packs

L_TOK = 361      # 19x19 feature map of a 300x300 frame


def trunk_hw(H, W):
    """Feature-map size of the ResNet-50-DC5 trunk (stride 16: stem conv s2, max-pool s2, layer2 s2,
    layer3 s2, layer4 dilated) for an HxW frame."""
    def down(n, k, p):
        return (n + 2 * p - k) // 2 + 1
    out = []
    for n in (H, W):
        n = down(n, 7, 3)       # conv1 7x7 s2 p3
        n = down(n, 3, 1)       # max-pool 3x3 s2 p1
        n = down(n, 3, 1)       # layer2 (3x3 s2 p1; the 1x1 s2 shortcut gives the same size)
        n = down(n, 3, 1)       # layer3
        out.append(n)
    return tuple(out)


_NEAREST = {}


def sample_masks_host(masks):
    """The padding mask at feature-map resolution, taken on the HOST from a CPU `data["masks"]`
    ([..., H, W], any dtype; nonzero = padded) -> uint8 [..., h, w].  The reference interpolates the
    full-resolution mask to the feature map with mode="nearest" on the device
    (detr_models/backbone.py:77); nearest picks one source pixel per output pixel, so only those
    h*w pixels have to cross PCIe (300x300 int64 frames: 720 KB -> 361 B each).  The source indices are
    produced by F.interpolate itself on an index ramp, so the sampling rule is torch's by construction."""
    H, W = masks.shape[-2:]
    h, w = trunk_hw(H, W)
    key = (H, W)
    if key not in _NEAREST:
        F = torch.nn.functional
        iy = F.interpolate(torch.arange(H, dtype=torch.float32)[None, None], size=h)[0, 0].long()
        ix = F.interpolate(torch.arange(W, dtype=torch.float32)[None, None], size=w)[0, 0].long()
        _NEAREST[key] = (iy, ix)
    iy, ix = _NEAREST[key]
    lead = masks.shape[:-2]
    m = masks.reshape(-1, H, W).index_select(1, iy).index_select(2, ix)
    return m.ne(0).to(torch.uint8).reshape(*lead, h, w)


class InnerLoop:
    def __init__(self, ops, detector, fusion_mod, kind, lr, clip=0.01):
        assert kind in ("A", "B") and (fusion_mod is not None or kind == "B")
        self.ops, self.detector, self.fusion_mod, self.kind = ops, detector, fusion_mod, kind
        self.lr, self.clip = float(lr), float(clip)
        self.theta_pack, self.theta_params, self.psi_pack, self.psi_params = detector_packs(detector)
        phi = list(fusion_mod.named_parameters()) if fusion_mod is not None else []
        self.phi_pack, self.phi_params = ParamPack(phi), [p for _, p in phi]
        # "gemm": the trunk on our own im2col + tf32x3 GEMM kernels (default); "cudnn": cuDNN convolutions
        self.backbone_impl = os.environ.get("ITN_BACKBONE", "gemm")
        self.backbone_tf32 = False      # cuDNN executor only: TF32 convs break the 1e-3 parity bar
        self.theta = self.psi = self.phi = None
        self.theta_t = self.psi_t = self.phi_t = None
        self._idx_cache = {}
        # train()-mode dropout (layers.DropCtx): DETR's 0.1 is hard-coded in the reference
        # (models/detr_models/transformer.py:278, new_transformer.py:23); fusion A reads its three from the YAML
        self.drop_p = 0.1
        self.fusion_drop = (0.1, 0.1, 0.1)          # residual, attention, embedding
        self.drop_seed = None                        # device int64[1], rewritten by the model before every train() step
        self.refresh_weights()

    # ------------------------------------------------------------------ weights
    def refresh_weights(self, repack=True):
        """(Re)pack theta / psi / phi into the persistent flat fp32 buffers (+ TF32-rounded twins
        in single-pass mode).  In place, so captured CUDA graphs keep reading valid addresses.
        repack=False: the flat buffers already hold the current weights (the Parameters alias them,
        trainer.MetaTrainerStep) - only the derived twins are rebuilt."""
        ops = self.ops
        dev = ops.device

        def fill(buf, pack, params):
            if buf is not None and not repack:
                return buf
            flat = pack.pack(params, device=dev, dtype=getattr(ops, "dtype", torch.float32))
            if buf is None:
                return flat.unsqueeze(0)
            buf[0].copy_(flat)
            return buf

        self.theta = fill(self.theta, self.theta_pack, self.theta_params)
        self.psi = fill(self.psi, self.psi_pack, self.psi_params)
        if not ops._clean:
            self.theta_r, self.psi_r = self.theta, self.psi
        else:
            self.theta_r = ops.round_tf32(self.theta, out=getattr(self, "theta_r", None))
            self.psi_r = ops.round_tf32(self.psi, out=getattr(self, "psi_r", None))
        if self.phi_params:
            self.phi = fill(self.phi, self.phi_pack, self.phi_params)
            self.phi_r = self.phi if not ops._clean else ops.round_tf32(self.phi, out=getattr(self, "phi_r", None))
        # W^T twins of the shared weights for the data-gradient GEMMs (K-major operands are ~4x
        # faster than MN-major ones through the fp32/tf32 tensor-core path)
        for nm, pack in (("theta", self.theta_pack), ("psi", self.psi_pack), ("phi", self.phi_pack)):
            src = getattr(self, nm + "_r", None) if getattr(self, nm, None) is not None else None
            if src is None:
                continue
            dst = getattr(self, nm + "_t", None)
            if dst is None:
                dst = ops.zeros(*src.shape)
                setattr(self, nm + "_t", dst)
            pack.transpose_into(ops, src, dst)
        # static weights of the tf32x3 GEMMs: keep their tf32 residuals beside them (ops.register_presplit)
        if hasattr(ops, "register_presplit"):
            for nm in ("theta", "psi", "phi", "theta_t", "psi_t", "phi_t"):
                buf = getattr(self, nm, None)
                if buf is not None:
                    ops.register_presplit(buf)

    def _det_weights(self, theta, theta_r, theta_t=None):
        return Weights((self.theta_pack, theta, theta_r, theta_t), (self.psi_pack, self.psi, self.psi_r, self.psi_t))

    def _fusion_weights(self):
        return Weights((self.phi_pack, self.phi, self.phi_r, self.phi_t))

    # ------------------------------------------------------------------ pieces
    def trunk(self, frames):
        """The frozen backbone on frames [N,3,H,W] -> channels-last features [N,h,w,2048]."""
        body = self.detector.backbone[0].body
        if self.backbone_impl == "gemm":
            return run_backbone_gemm(body, frames, self.ops)
        return run_backbone(body, frames, self.backbone_tf32)

    def features(self, frames, masks, src=None):
        """frames [N,3,H,W], masks [N,H,W] (nonzero = padded) -> (src_r [N,L,2048] TF32-clean
        token-major, pos [N*L,256], kmask uint8 [N,L], hw).  self.src keeps the unrounded features.
        src: trunk features [N,h,w,2048] computed by the caller (graph.PipelinedPredict runs the trunk chunk by
        chunk while the rest of the frames is still crossing PCIe); `frames` is then not read."""
        ops = self.ops
        if src is None:
            src = self.trunk(frames)                                                      # [N,h,w,2048]
        N, h, w, C = src.shape
        self.src = src.reshape(N, h * w, C)
        src_r = ops.round_tf32(self.src)
        if masks.dtype == torch.uint8 and tuple(masks.shape[-2:]) == (h, w):
            m = masks.to(torch.bool)               # sampled on the host already (sample_masks_host)
        else:
            m = torch.nn.functional.interpolate(masks[None].float(), size=(h, w)).to(torch.bool)[0]
        pos = ops.pos_embed_sine(m).reshape(N * h * w, -1)
        kmask = m.reshape(N, h * w).to(torch.uint8).contiguous()
        return src_r, pos, kmask, (h, w)

    def detect(self, frames, masks, want_preds=False):
        """Plain DETR forward with theta on N frames (no adaptation).  want_preds: also build the
        TF32-clean prediction tokens [N*50, 1496] that fusion embeds (returned as out["preds"])."""
        N = frames.shape[0]
        src_r, pos, kmask, hw = self.features(frames, masks)
        L = hw[0] * hw[1]
        W = self._det_weights(self.theta, self.theta_r)
        C = self.detector.class_embed.out_features
        preds = self.ops.empty(N * detr_t.NQ, detr_t.D + C + 4) if want_preds else None
        out, _ = detr_t.detr_t_forward(self.ops, W, src_r.view(1, N * L, -1), pos, kmask, 1, N, L,
                                       preds=preds, need_cache=False)
        out["preds"] = preds
        return out, src_r, hw

    # dropout contexts of one step: every forward pass draws from its own range of sites
    PASS_SITES = {"pre": 0, "fusion": 2048, "post": 4096, "post1": 6144}

    def drop_ctx(self, which, train):
        """layers.DropCtx of pass `which` ("pre", "fusion", "post", "post1") in train() mode, None in eval()."""
        if not train:
            return None
        from .layers import DropCtx
        if self.drop_seed is None:
            raise RuntimeError("train()-mode pass without a dropout seed: call it through the model")
        if which == "fusion":
            r, a, e = self.fusion_drop
            return DropCtx(r, self.drop_seed, self.PASS_SITES[which], p_attn=a, p_embd=e)
        return DropCtx(self.drop_p, self.drop_seed, self.PASS_SITES[which])

    # ------------------------------------------------------------------ the hot path
    def adapt_detect(self, frames, masks, post_frames=(0,), want_trace=False, train=False, src=None):
        """frames [E,S,3,H,W], masks [E,S,H,W] on the device -> dict with post-adapt
        pred_logits [E,P,50,C], pred_boxes [E,P,50,4] (P = len(post_frames)) and features.
        train: the reference's train() mode - dropout in the pre-adapt pass, the fusion network and the
        post-adapt pass (models/interactron.py:31-59 run under model.train())."""
        ops = self.ops
        E, S = frames.shape[:2]
        src_r, pos, kmask, (h, w) = self.features(frames.flatten(0, 1), masks.flatten(0, 1), src=src)
        L = h * w
        C = self.detector.class_embed.out_features
        # -- pre-adapt pass (shared theta) and learned loss
        Wd = self._det_weights(self.theta, self.theta_r, self.theta_t)
        preds = ops.empty(E * S * detr_t.NQ, detr_t.D + C + 4)
        pre, cache = detr_t.detr_t_forward(ops, Wd, src_r.view(E, S * L, -1), pos, kmask, E, S, L, preds=preds,
                                           drop=self.drop_ctx("pre", train))
        Wf = self._fusion_weights()
        if self.kind == "A":
            fout, fcache = fusion.fusion_a_forward(ops, Wf, pre["memory_r"], preds, E, S, L,
                                                   drop=self.drop_ctx("fusion", train))
            dmemory, dpreds = fusion.fusion_a_backward(ops, Wf, fcache)
        else:
            fout, fcache = fusion.fusion_b_forward(ops, Wf, pre["memory_r"], preds, E, S, L,
                                                   drop=self.drop_ctx("fusion", train))
            dmemory, dpreds = fusion.fusion_b_backward(ops, Wf, fcache)
        # -- inner gradient g_e = d learned_loss_e / d theta
        g = ops.empty(E, self.theta_pack.numel)
        detr_t.detr_t_backward(ops, Wd, cache, GradSink(ops, self.theta_pack, g), dpreds=dpreds, dmemory=dmemory)
        del cache, fcache
        # -- fast weights
        theta_p, theta_p_r = ops.sgd_clip_update(self.theta, g, self.lr, self.clip)
        # -- post-adapt pass on the requested frames with per-episode weights
        P = len(post_frames)
        if P == S and tuple(post_frames) == tuple(range(S)):
            src_p, pos_p, km_p, src_full = src_r.view(E, S * L, -1), pos, kmask, self.src
        else:
            key = (E, S, tuple(post_frames))
            if key not in self._idx_cache:          # built once, outside any graph capture
                self._idx_cache[key] = torch.tensor([e * S + f for e in range(E) for f in post_frames],
                                                    device=src_r.device)
            idx = self._idx_cache[key]
            src_p = src_r.index_select(0, idx).view(E, P * L, -1)
            src_full = self.src.index_select(0, idx)
            pos_p = pos.view(E * S, L, -1).index_select(0, idx).reshape(E * P * L, -1)
            km_p = kmask.index_select(0, idx)
        Wp = self._det_weights(theta_p, theta_p_r)
        post, _ = detr_t.detr_t_forward(ops, Wp, src_p, pos_p, km_p, E, P, L, need_cache=False,
                                        drop=self.drop_ctx("post", train))
        out = {
            "pred_logits": post["logits"].view(E, P, detr_t.NQ, C),
            "pred_boxes": post["boxes"].view(E, P, detr_t.NQ, 4),
            "box_features": post["hs"].view(E, P, detr_t.NQ, detr_t.D),
            "embedded_memory_features": post["memory"].view(E, P, h, w, detr_t.D).permute(0, 1, 4, 2, 3),
            "image_features": src_full.view(E, P, h, w, -1).permute(0, 1, 4, 2, 3),
            "learned_loss": fout["learned_loss"],
            "actions": fout["actions"],
        }
        if want_trace:
            out["trace"] = dict(pre_logits=pre["logits"].view(E, S, detr_t.NQ, C),
                                pre_boxes=pre["boxes"].view(E, S, detr_t.NQ, 4),
                                loss_vec=fout["loss_vec"], g=g, theta_prime=theta_p)
        return out
