"""Attention-score GEMM in isolation (QK^T of the DETR encoder, 32 episodes): [B*H, Lq, hd] x [B*H, hd, Lk] into
padded score rows, as layers.attention_fwd launches it.  python tools/gemm_scores.py [bn] [Lq Lk hd batch]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] != "0":
    os.environ["ITN_GEMM_BN"] = sys.argv[1]
from interactron_b200.ops import CudaOps  # noqa: E402

Lq, Lk, hd, B = (int(x) for x in sys.argv[2:6]) if len(sys.argv) > 5 else (361, 361, 32, 160)
nh = 8
ops = CudaOps()
q = torch.randn(B, Lq, nh * hd, device="cuda")
k = torch.randn(B, Lk, nh * hd, device="cuda")
qh = q.reshape(B, Lq, nh, hd).permute(0, 2, 1, 3)
khT = k.reshape(B, Lk, nh, hd).permute(0, 2, 3, 1)
ld = (Lk + 3) // 4 * 4
P = torch.empty(B, nh, Lq, ld, device="cuda")
p = P[..., :Lk]
for _ in range(3):
    ops.matmul(qh, khT, out=p, out_pad=True)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(10):
        ops.matmul(qh, khT, out=p, out_pad=True)
g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
g.replay()
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 100
print(f"scores {B}x{nh} [{Lq}x{hd}]x[{hd}x{Lk}] BN={os.environ.get('ITN_GEMM_BN', 'auto')} DBG={os.environ.get('ITN_GEMM_DBG', '0')}: "
      f"{us:.1f} us  {P.numel() * 4 / us / 1e6:.2f} TB/s written  {2.0 * B * nh * Lq * Lk * hd / us / 1e6:.1f} TF/s")
