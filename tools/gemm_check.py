"""GPU diagnostic for itn_gemm_tf32 / itn_gemm_simt: accuracy table vs fp64 torch and
rough throughput.  Each group of cases runs in its own subprocess under a timeout so a
hung kernel cannot eat the whole GPU lease.

    python tools/gemm_check.py            # driver: all groups
    python tools/gemm_check.py --group kk # one group in-process
"""
import argparse
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

GROUPS = ["kk", "km", "mk", "mm", "batch", "epi", "perf"]


def rel_err(out, ref):
    return ((out.double() - ref).norm() / ref.norm().clamp_min(1e-30)).item()


def run_group(group):
    import torch
    from interactron_b200.ops import CudaOps

    ops = CudaOps()
    torch.manual_seed(0)
    dev = "cuda"

    def mk(shape, transposed):
        """Logical [..., R, Cc] tensor, stored transposed if requested."""
        if transposed:
            t = torch.randn(*shape[:-2], shape[-1], shape[-2], device=dev)
            return t.transpose(-1, -2)
        return torch.randn(*shape, device=dev)

    def case(name, M, N, K, a_mn, b_mn, batch=(), **kw):
        a = mk((*batch, M, K), a_mn)          # A K-major unless a_mn
        b = mk((*batch, K, N), not b_mn)      # B K-major (stored [N,K]) unless b_mn
        ref = a.double() @ b.double()
        res = []
        for impl in ("tf32", "simt"):
            ops.force_simt = impl == "simt"
            try:
                n0 = ops.n_tf32
                out = ops.matmul(a, b, **kw)
                torch.cuda.synchronize()
                tag = impl if (impl == "simt" or ops.n_tf32 > n0) else "tf32->simt"
                res.append(f"{tag}={rel_err(out, ref):.2e}")
            except Exception as e:  # noqa
                res.append(f"{impl}=ERR({str(e)[:80]})")
        ops.force_simt = False
        print(f"  {name:34s} M={M:5d} N={N:5d} K={K:5d} batch={batch!s:8s} " + " ".join(res), flush=True)

    shapes = [(128, 128, 32), (128, 128, 256), (256, 256, 64), (1805, 256, 256), (250, 1236, 256),
              (250, 512, 1496), (361, 32, 361), (50, 361, 32), (130, 70, 40), (2060, 2048, 512),
              (250, 4, 256), (250, 256, 4), (256, 2048, 1805), (1236, 256, 250), (364, 32, 361),
              (132, 68, 40), (512, 1496, 250)]
    if group in ("kk", "km", "mk", "mm"):
        a_mn, b_mn = group[0] == "m", group[1] == "m"
        for bn in ("0", "32", "64", "128", "256"):
            os.environ["ITN_GEMM_BN"] = bn
            print(f" [BN override {bn}]")
            for (M, N, K) in shapes:
                # MN-major operands need the contiguous dim to be a multiple of 4 for TMA
                case(f"{group}", M, N, K, a_mn, b_mn)
        os.environ["ITN_GEMM_BN"] = "0"
    elif group == "batch":
        case("batch b0", 361, 361, 32, False, False, batch=(40,))
        case("batch b0,b1", 361, 32, 361, False, True, batch=(5, 8))
        case("batch mk", 364, 32, 361, True, True, batch=(5, 8))
        # broadcast weight over batch
        import torch
        a = torch.randn(6, 300, 256, device=dev)
        w = torch.randn(512, 256, device=dev)
        out = ops.matmul(a, w.t())
        ref = a.double() @ w.double().t()
        print(f"  broadcast B          {rel_err(out, ref):.2e}")
        # heads view: q [F, L, H, hd] -> [F, H, L, hd]
        F_, L, H, hd = 5, 361, 8, 32
        q = torch.randn(F_, L, H * hd, device=dev)
        k = torch.randn(F_, L, H * hd, device=dev)
        v = torch.randn(F_, L, H * hd, device=dev)
        qh = q.view(F_, L, H, hd).permute(0, 2, 1, 3)
        kh = k.view(F_, L, H, hd).permute(0, 2, 1, 3)
        vh = v.view(F_, L, H, hd).permute(0, 2, 1, 3)
        ldp = (L + 3) // 4 * 4
        sbuf = torch.zeros(F_, H, L, ldp, device=dev)
        s = sbuf[..., :L]
        ops.matmul(qh, kh.transpose(-1, -2), out=s)
        ref = qh.double() @ kh.double().transpose(-1, -2)
        print(f"  heads QK^T           {rel_err(s, ref):.2e}")
        o = torch.empty(F_, L, H * hd, device=dev)
        oh = o.view(F_, L, H, hd).permute(0, 2, 1, 3)
        ops.matmul(s, vh, out=oh)
        ref2 = s.double() @ vh.double()
        print(f"  heads PV (strided C) {rel_err(oh, ref2):.2e}")
        dv = ops.matmul(s.transpose(-1, -2), oh)
        ref3 = s.double().transpose(-1, -2) @ oh.double()
        print(f"  heads P^T dO         {rel_err(dv, ref3):.2e}")
    elif group == "epi":
        import torch
        M, N, K = 300, 260, 96
        a = torch.randn(M, K, device=dev)
        w = torch.randn(N, K, device=dev)
        bias = torch.randn(N, device=dev)
        res = torch.randn(M, N, device=dev)
        aux = torch.randn(M, N, device=dev)
        base = a.double() @ w.double().t()
        for impl in ("tf32", "simt"):
            ops.force_simt = impl == "simt"
            o = ops.matmul(a, w.t(), bias=bias)
            print(f"  {impl} bias        {rel_err(o, base + bias.double()):.2e}")
            o = ops.matmul(a, w.t(), bias=bias, act="relu", residual=res)
            print(f"  {impl} relu+res    {rel_err(o, torch.relu(base + bias.double()) + res.double()):.2e}")
            pre = torch.empty(M, N, device=dev)
            o = ops.matmul(a, w.t(), bias=bias, act="gelu", out_pre=pre)
            print(f"  {impl} gelu        {rel_err(o, torch.nn.functional.gelu(base + bias.double())):.2e}"
                  f"  pre {rel_err(pre, base + bias.double()):.2e}")
            o = ops.matmul(a, w.t(), epi="relu_mask", aux=aux)
            print(f"  {impl} relu_mask   {rel_err(o, base * (aux > 0).double()):.2e}")
            x = aux.double().requires_grad_(True)
            torch.nn.functional.gelu(x).sum().backward()
            o = ops.matmul(a, w.t(), epi="gelu_grad", aux=aux)
            print(f"  {impl} gelu_grad   {rel_err(o, base * x.grad):.2e}")
            c = res.clone()
            ops.matmul(a, w.t(), out=c, accumulate=True, alpha=0.5)
            print(f"  {impl} accumulate  {rel_err(c, 0.5 * base + res.double()):.2e}")
        ops.force_simt = False
    elif group == "perf":
        import torch
        for (M, N, K, batch) in [(8 * 1805, 256, 2048, ()), (8 * 1805, 2048, 256, ()), (8 * 1805, 256, 256, ()),
                                 (8 * 2060, 2048, 512, ()), (8192, 8192, 2048, ()), (361, 361, 32, (320,)),
                                 (2060, 2060, 64, (64,)), (2000, 512, 512, ())]:
            for bn, prec in (("128", "tf32"), ("256", "tf32"), ("64", "tf32x3"), ("128", "tf32x3"), ("256", "tf32x3")):
                os.environ["ITN_GEMM_BN"] = bn
                ops.precision = prec
                a = torch.randn(*batch, M, K, device=dev)
                w = torch.randn(*batch, N, K, device=dev)
                out = torch.empty(*batch, M, N, device=dev)
                for _ in range(3):
                    ops.matmul(a, w.transpose(-1, -2), out=out)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                iters = 20
                e0.record()
                for _ in range(iters):
                    ops.matmul(a, w.transpose(-1, -2), out=out)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / iters
                nb = 1
                for x in batch:
                    nb *= x
                tf = 2.0 * M * N * K * nb / ms / 1e9
                # cuBLAS TF32 comparison (library baseline)
                torch.backends.cuda.matmul.allow_tf32 = True
                for _ in range(3):
                    torch.matmul(a, w.transpose(-1, -2), out=out)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(iters):
                    torch.matmul(a, w.transpose(-1, -2), out=out)
                e1.record()
                torch.cuda.synchronize()
                ms2 = e0.elapsed_time(e1) / iters
                tf2 = 2.0 * M * N * K * nb / ms2 / 1e9
                print(f"  perf M={M} N={N} K={K} batch={batch} BN={bn} {prec:6s}: {ms*1e3:8.1f} us {tf:8.1f} TF/s"
                      f"   | cuBLAS tf32 {ms2*1e3:8.1f} us {tf2:8.1f} TF/s", flush=True)
        os.environ["ITN_GEMM_BN"] = "0"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--group", default=None)
    ap.add_argument("--timeout", type=int, default=300)
    args = ap.parse_args()
    if args.group:
        run_group(args.group)
        return
    import torch  # noqa: F401  (warm the page cache before the timed subprocesses)
    for g in GROUPS:
        print(f"== group {g}", flush=True)
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--group", g], timeout=args.timeout,
                               capture_output=True, text=True)
            print(p.stdout, end="")
            if p.returncode != 0:
                print(f"  !! exit {p.returncode}\n{p.stderr[-1500:]}")
        except subprocess.TimeoutExpired as e:
            out = e.stdout.decode() if isinstance(e.stdout, bytes) else (e.stdout or "")
            print(out, end="")
            print(f"  !! TIMEOUT after {args.timeout}s (hung kernel?)")
        print(f"   ({time.time()-t0:.1f}s)", flush=True)


if __name__ == "__main__":
    main()
