"""Implicit-GEMM convolution (TMA im2col tensor map inside itn_gemm_tf32) against the explicit im2col + GEMM
path and torch's fp64 conv2d, on the geometries of the ResNet-50 trunk; with error localisation and timing."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interactron_b200.ops import CudaOps  # noqa: E402


def run(ops, N, H, W, Cin, Cout, k, stride, pad, dil, time_it=False):
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(N, H, W, Cin, device="cuda", generator=g)
    w = torch.randn(Cout, k, k, Cin, device="cuda", generator=g) * (k * k * Cin) ** -0.5
    b = torch.randn(Cout, device="cuda", generator=g)
    wm = w.reshape(Cout, -1).contiguous()
    y, Ho, Wo = ops.conv_gemm(x, wm, k, k, stride, pad, dil, bias=b, act="relu")
    a, Ho2, Wo2 = ops.im2col_nhwc(x, k, k, stride, pad, dil)
    y2 = ops.matmul(a, wm.t(), bias=b, act="relu")
    ref = F.relu(F.conv2d(x.permute(0, 3, 1, 2).double(), w.permute(0, 3, 1, 2).double(), b.double(), stride, pad, dil))
    ref = ref.permute(0, 2, 3, 1).reshape(-1, Cout)
    torch.cuda.synchronize()
    e = ((y.double() - ref).norm() / ref.norm()).item()
    e2 = ((y2.double() - ref).norm() / ref.norm()).item()
    same = torch.equal(y, y2)
    msg = f"N={N} {H}x{W} Cin={Cin} Cout={Cout} k={k} s={stride} p={pad} d={dil} -> {Ho}x{Wo}: implicit {e:.2e} explicit {e2:.2e} bit-equal {same}"
    if e > 1e-4:
        err = (y.double() - ref).view(N, Ho, Wo, Cout).pow(2).sum(-1).sqrt()
        nrm = ref.view(N, Ho, Wo, Cout).pow(2).sum(-1).sqrt()
        bad = (err > 1e-3 * nrm.clamp_min(1e-9))
        print("[FAIL]", msg)
        print("   bad pixels per image:", bad.flatten(1).sum(1).tolist()[:8], "of", Ho * Wo)
        print("   bad rows (py) of image 0:", bad[0].any(1).nonzero().flatten().tolist()[:40])
        print("   bad cols (px) of image 0:", bad[0].any(0).nonzero().flatten().tolist()[:40])
        return False
    print("[ok]  ", msg, flush=True)
    if time_it:
        def t(fn, n=5):
            fn(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                fn()
            e1.record(); torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n * 1e3
        ti = t(lambda: ops.conv_gemm(x, wm, k, k, stride, pad, dil, bias=b, act="relu"))
        tc = t(lambda: ops.im2col_nhwc(x, k, k, stride, pad, dil))
        tg = t(lambda: ops.matmul(a, wm.t(), bias=b, act="relu"))
        print(f"      implicit {ti:8.1f} us | im2col {tc:8.1f} + gemm {tg:8.1f} = {tc + tg:8.1f} us", flush=True)
    return True


def main():
    ops = CudaOps()
    ok = True
    ok &= run(ops, 1, 8, 8, 32, 32, 3, 1, 1, 1)
    ok &= run(ops, 2, 19, 19, 64, 64, 3, 1, 1, 1)
    ok &= run(ops, 2, 19, 19, 64, 64, 3, 1, 2, 2)          # dilated (layer4)
    ok &= run(ops, 3, 38, 38, 64, 128, 3, 2, 1, 1)         # strided 3x3 (layer2/3 first block)
    ok &= run(ops, 3, 38, 38, 64, 128, 1, 2, 0, 1)         # strided 1x1 (downsample)
    ok &= run(ops, 2, 75, 75, 64, 64, 3, 1, 1, 1)
    ok &= run(ops, 1, 16, 16, 4, 64, 3, 1, 1, 1)           # 4-channel mode: 9 taps = 2 k-blocks
    ok &= run(ops, 2, 30, 30, 4, 64, 7, 2, 3, 1)           # the stem's geometry, small
    ok &= run(ops, 3, 300, 300, 4, 64, 7, 2, 3, 1)         # the RGB stem (zero-padded to 4 channels)
    if len(sys.argv) > 1:
        return 0 if ok else 1
    E = 160
    run(ops, 310, 300, 300, 4, 64, 7, 2, 3, 1, True)        # stem at 62 episodes
    run(ops, E, 75, 75, 64, 64, 3, 1, 1, 1, True)           # layer1 conv2
    run(ops, E, 75, 75, 128, 128, 3, 2, 1, 1, True)         # layer2.0 conv2
    run(ops, E, 38, 38, 128, 128, 3, 1, 1, 1, True)         # layer2 conv2
    run(ops, E, 38, 38, 256, 256, 3, 2, 1, 1, True)         # layer3.0 conv2
    run(ops, E, 19, 19, 256, 256, 3, 1, 1, 1, True)         # layer3 conv2
    run(ops, E, 19, 19, 512, 512, 3, 1, 2, 2, True)         # layer4 conv2 (dilated)
    run(ops, E, 75, 75, 256, 512, 1, 2, 0, 1, True)         # layer2 downsample
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
