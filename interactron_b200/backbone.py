"""Frozen ResNet-50-DC5 trunk (decision D1, SURVEY.md section 8a row D1: the backbone is not a fast
weight, so its features are computed once per episode and reused by the pre- and post-adaptation
passes; reference models/detr_models/backbone.py:57-92 + torchvision resnet50).

Because the trunk is frozen, the FrozenBatchNorm affine of every convolution is folded into the
convolution's weight and bias once (the reference applies it as separate element-wise ops,
backbone.py:44-54).  Two executors:

  GemmTrunk   (default) channels-last activations; every convolution is ONE tf32x3 tensor-core GEMM
              of our own kernel with bias / ReLU / residual fused in the epilogue: 1x1 convolutions
              read the activation matrix [N*H*W, Cin] directly, kxk ones (and the strided 1x1
              down-samples) go through a channels-last im2col gather first.  fp32-accurate
              (error-compensated TF32), ~3-4x faster than cuDNN's fp32 convolutions on B200.
  FoldedTrunk cuDNN fused ConvolutionBiasActivation in strict fp32 (NCHW) or TF32 (channels-last);
              kept as a cross-check (`ITN_BACKBONE=cudnn`) and for the TF32-backbone ablation.
"""
import torch
import torch.nn.functional as F


class FoldedTrunk:
    """BN-folded weights of the ResNet body + the fused forward."""

    def __init__(self, body):
        self.body = body
        self.key = None
        self.layers = None

    def _key(self):
        return tuple((t.data_ptr(), t._version) for t in list(self.body.parameters()) + list(self.body.buffers()))

    @staticmethod
    def _fold(conv, bn):
        scale, shift = bn.scale_shift()
        return (conv.weight.detach() * scale.view(-1, 1, 1, 1)).contiguous(), shift.detach().contiguous()

    def refresh(self):
        key = self._key()
        if key == self.key:
            return
        b = self.body
        with torch.no_grad():
            stem = self._fold(b.conv1, b.bn1)
            blocks = []
            for name in ("layer1", "layer2", "layer3", "layer4"):
                for blk in getattr(b, name):
                    d = None
                    if blk.downsample is not None:
                        d = self._fold(blk.downsample[0], blk.downsample[1]) + (blk.downsample[0].stride,)
                    blocks.append(dict(
                        a=self._fold(blk.conv1, blk.bn1), b=self._fold(blk.conv2, blk.bn2),
                        c=self._fold(blk.conv3, blk.bn3), stride=blk.conv2.stride, pad=blk.conv2.padding,
                        dil=blk.conv2.dilation, down=d))
        self.layers = (stem, blocks)
        self.key = key

    @staticmethod
    def _conv_relu(x, w, b, stride=(1, 1), pad=(0, 0), dil=(1, 1)):
        if x.is_cuda:
            return torch.cudnn_convolution_relu(x, w, b, stride, pad, dil, 1)
        return F.relu(F.conv2d(x, w, b, stride, pad, dil))

    @staticmethod
    def _conv_add_relu(x, w, b, z):
        if x.is_cuda:
            return torch.cudnn_convolution_add_relu(x, w, z, 1.0, b, (1, 1), (0, 0), (1, 1), 1)
        return F.relu(F.conv2d(x, w, b) + z)

    def forward(self, x):
        self.refresh()
        stem, blocks = self.layers
        x = self._conv_relu(x, stem[0], stem[1], (2, 2), (3, 3))
        x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
        for blk in blocks:
            idt = x
            if blk["down"] is not None:
                wd, bd, sd = blk["down"]
                idt = F.conv2d(x, wd, bd, sd)
            o = self._conv_relu(x, *blk["a"])
            o = self._conv_relu(o, blk["b"][0], blk["b"][1], blk["stride"], blk["pad"], blk["dil"])
            x = self._conv_add_relu(o, blk["c"][0], blk["c"][1], idt)
        return x


class GemmTrunk:
    """The trunk as im2col + GEMM on the hand-written kernels (see module docstring)."""

    def __init__(self, body, ops):
        self.body, self.ops = body, ops
        self.key = None
        self.layers = None

    _key = FoldedTrunk._key

    def _prep(self, conv, bn, pad_cin=0):
        """-> dict(w [Cout, ld] with columns (ky, kx, cin) matching im2col, bias, geometry).
        pad_cin: zero input channels appended (the RGB stem runs with 4 channels so that the
        im2col gather moves 16-byte vectors)."""
        w, b = FoldedTrunk._fold(conv, bn)
        if pad_cin:
            w = F.pad(w, (0, 0, 0, 0, 0, pad_cin))
        co, ci, kh, kw = w.shape
        k = kh * kw * ci
        ld = (k + 3) // 4 * 4
        wm = torch.zeros(co, ld, dtype=w.dtype, device=w.device)
        wm[:, :k] = w.permute(0, 2, 3, 1).reshape(co, k)
        wm = self.ops.round_tf32(wm.to(self.ops.device).contiguous())
        return dict(w=wm, b=b.to(self.ops.device), kh=kh, kw=kw, stride=conv.stride[0], pad=conv.padding[0],
                    dil=conv.dilation[0], cout=co)

    def refresh(self):
        key = self._key()
        if key == self.key:
            return
        b = self.body
        with torch.no_grad():
            stem = self._prep(b.conv1, b.bn1, pad_cin=1)
            blocks = []
            for name in ("layer1", "layer2", "layer3", "layer4"):
                for blk in getattr(b, name):
                    blocks.append(dict(
                        a=self._prep(blk.conv1, blk.bn1), b=self._prep(blk.conv2, blk.bn2),
                        c=self._prep(blk.conv3, blk.bn3),
                        down=self._prep(blk.downsample[0], blk.downsample[1]) if blk.downsample is not None else None))
        self.layers = (stem, blocks)
        self.key = key
        if hasattr(self.ops, "register_presplit"):          # frozen trunk: every conv weight is a static B operand
            for L in [stem] + [blk[k] for blk in blocks for k in ("a", "b", "c", "down") if blk[k] is not None]:
                self.ops.register_presplit(L["w"])

    def _conv(self, x, L, act="relu", residual=None):
        """x [N,H,W,Cin] channels-last -> [N,Ho,Wo,Cout]; bias (+ residual) + ReLU fused in the GEMM."""
        ops = self.ops
        N, H, W, Cin = x.shape
        if L["kh"] == 1 and L["stride"] == 1:
            a, Ho, Wo = x.view(N * H * W, Cin), H, W
        elif ((Cin % 32 == 0 or (Cin == 4 and getattr(ops, "implicit_stem", False))) and getattr(ops, "implicit_conv", False)
              and not ops._clean and not ops.force_simt):
            # 3x3 / strided convolutions: the TMA unit gathers the patches (im2col tensor map), nothing is materialised
            res = residual.reshape(-1, L["cout"]) if residual is not None else None
            y, Ho, Wo = ops.conv_gemm(x, L["w"], L["kh"], L["kw"], L["stride"], L["pad"], L["dil"], bias=L["b"],
                                      act=act, residual=res, act_after_residual=res is not None)
            return y.view(N, Ho, Wo, L["cout"])
        else:
            a, Ho, Wo = ops.im2col_nhwc(x, L["kh"], L["kw"], L["stride"], L["pad"], L["dil"])
        res = residual.view(N * Ho * Wo, L["cout"]) if residual is not None else None
        y = ops.matmul(a, L["w"].t(), bias=L["b"], act=act, residual=res, act_after_residual=res is not None)
        return y.view(N, Ho, Wo, L["cout"])

    def forward(self, frames):
        """frames [N,3,H,W] -> [N,h,w,2048] channels-last features."""
        self.refresh()
        stem, blocks = self.layers
        x = F.pad(frames.permute(0, 2, 3, 1), (0, 1)).contiguous()      # NHWC, RGB + one zero channel
        x = self._conv(x, stem)
        x = self.ops.maxpool3x3s2_nhwc(x)
        for blk in blocks:
            idt = x if blk["down"] is None else self._conv(x, blk["down"], act=None)
            o = self._conv(x, blk["a"])
            o = self._conv(o, blk["b"])
            x = self._conv(o, blk["c"], residual=idt)
        return x


_TRUNKS = {}


def run_backbone_gemm(body, frames, ops):
    """The default executor: our own kernels end to end (see GemmTrunk)."""
    key = (id(body), id(ops))
    trunk = _TRUNKS.get(key)
    if trunk is None or trunk.body is not body:
        trunk = _TRUNKS[key] = GemmTrunk(body, ops)
    with torch.no_grad():
        return trunk.forward(frames)


def run_backbone(body, frames, tf32=False):
    """cuDNN executor.  frames [N,3,H,W] -> token-major features [N,h,w,2048] (fp32, contiguous).
    tf32=False keeps cuDNN in strict fp32: with TF32 convolutions the features move by ~4e-4,
    which the adaptation step amplifies to ~1.6e-2 on the re-detected logits (measured,
    profiles/README.md) — outside the 1e-3 parity bar.  fp32 runs in NCHW: cuDNN's fp32
    channels-last path falls back to a 100 ms direct kernel for the two dilated 3x3 convolutions
    of layer4 (257 ms vs 35 ms per 40 frames, measured); TF32 is fastest channels-last."""
    trunk = _TRUNKS.get(id(body))
    if trunk is None or trunk.body is not body:
        trunk = _TRUNKS[id(body)] = FoldedTrunk(body)
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = bool(tf32)
    try:
        with torch.no_grad():
            x = frames.contiguous(memory_format=torch.channels_last) if tf32 else frames.contiguous()
            y = trunk.forward(x)
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    return y.permute(0, 2, 3, 1).contiguous()
