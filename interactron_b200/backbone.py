"""Frozen ResNet-50-DC5 trunk, executed through PyTorch/cuDNN (decision D1, SURVEY.md section 8a row D1:
the backbone is not a fast weight, so its features are computed once per episode and reused by
the pre- and post-adaptation passes).  This is the only arithmetic on the path that is a library
call; everything downstream of `src` runs in the hand-written kernels.
"""
import torch


def run_backbone(body, frames, tf32=False):
    """frames [N,3,H,W] -> token-major features [N,h,w,2048] (fp32, contiguous).
    tf32=False keeps cuDNN in strict fp32: with TF32 convolutions the features move by ~4e-4,
    which the adaptation step amplifies to ~1.6e-2 on the re-detected logits (measured,
    profiles/README.md) — outside the 1e-3 parity bar.  fp32 runs in NCHW: cuDNN's fp32
    channels-last path falls back to a 100 ms direct kernel for the two dilated 3x3 convolutions
    of layer4 (257 ms vs 35 ms per 40 frames, measured); TF32 is fastest channels-last."""
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = bool(tf32)
    try:
        with torch.no_grad():
            x = frames.contiguous(memory_format=torch.channels_last) if tf32 else frames.contiguous()
            y = body(x)["0"]
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    return y.permute(0, 2, 3, 1).contiguous()
