"""In-kernel timeline of the persistent GEMM (debug build `make -C interactron_b200/csrc trace`).
python tools/gemm_trace.py M N K [prec] [bn]   -> per k-block / per tile hand-off times of CTA 0 (SM cycles)."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from interactron_b200 import _lib  # noqa: E402

_lib.LIB_PATH = os.path.join(ROOT, "interactron_b200", "libinteractron_b200_trace.so")
from interactron_b200.ops import CudaOps  # noqa: E402

M, N, K = (int(x) for x in sys.argv[1:4])
prec = sys.argv[4] if len(sys.argv) > 4 else "tf32x3"
if len(sys.argv) > 5 and sys.argv[5] != "0":
    os.environ["ITN_GEMM_BN"] = sys.argv[5]
pad = len(sys.argv) > 6 and sys.argv[6] == "pad"     # rows padded to 4 columns, 128-bit stores (c_pad)
ld_out = int(sys.argv[7]) if len(sys.argv) > 7 else 0  # explicit row pitch of the output (floats)
ops = CudaOps()
ops.precision = prec
ops.lib.itn_debug_set_trace.argtypes = [C.c_void_p]
a = torch.randn(M, K, device="cuda")
w = torch.randn(N, K, device="cuda")
out = torch.empty(M, ld_out if ld_out else ((N + 3) // 4 * 4 if pad else N), device="cuda")[:, :N]
for _ in range(3):
    ops.matmul(a, w.t(), out=out, out_pad=pad)
torch.cuda.synchronize()
buf = torch.zeros(9, 1024, dtype=torch.int64, device="cuda")
assert ops.lib.itn_debug_set_trace(C.c_void_p(buf.data_ptr())) == 0
ops.matmul(a, w.t(), out=out, out_pad=pad)
torch.cuda.synchronize()
t = buf.cpu()
ph = t[8, :4].tolist()
marks = t[8, 8:11].tolist()
t[8] = 0
t0 = int(marks[0]) if marks[0] else int(t[t > 0].min())
print(f"CTA 0: entry 0, prologue done {marks[1] - t0}, all roles done {marks[2] - t0} (SM cycles)")
if ph[3]:
    print(f"epilogue warp 0 of CTA 0: {int(ph[3])} chunks; cycles per chunk: TMEM read {ph[0]/ph[3]:.0f}, "
          f"transpose {ph[1]/ph[3]:.0f}, math+stores {ph[2]/ph[3]:.0f}")
nkb = (K + 31) // 32
import signal
signal.signal(signal.SIGPIPE, signal.SIG_DFL)
names = ["tma_issue", "split_start", "split_done", "mma_ready", "mma_issued"]
print(f"M={M} N={N} K={K} {prec}: k-blocks per tile {nkb}; times in SM cycles since first event")
print("kb   " + " ".join(f"{n:>12s}" for n in names))
for i in range(min(6 * nkb, 40)):
    print(f"{i:3d}  " + " ".join(f"{int(t[s, i]) - t0 if t[s, i] > 0 else -1:12d}" for s in (0, 1, 2, 3, 4)))
print("tile  mma_tempty_ok   epi_tfull    epi_done")
for i in range(8):
    print(f"{i:3d}  " + " ".join(f"{int(t[s, i]) - t0 if t[s, i] > 0 else -1:12d}" for s in (7, 5, 6)))
