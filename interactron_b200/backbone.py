"""Frozen ResNet-50-DC5 trunk, executed through PyTorch/cuDNN (decision D1, SURVEY.md section 8a row D1:
the backbone is not a fast weight, so its features are computed once per episode and reused by
the pre- and post-adaptation passes).  This is the only arithmetic on the path that is a library
call; everything downstream of `src` runs in the hand-written kernels.
"""
import torch


def run_backbone(body, frames):
    """frames [N,3,H,W] -> channels-last features [N,h,w,2048] (fp32, contiguous)."""
    with torch.no_grad():
        x = frames.contiguous(memory_format=torch.channels_last)
        y = body(x)["0"]
    return y.permute(0, 2, 3, 1).contiguous()


def set_backbone_precision(tf32):
    """cuDNN convolutions in TF32 (default on Blackwell) or strict fp32 (parity tests)."""
    torch.backends.cudnn.allow_tf32 = bool(tf32)
