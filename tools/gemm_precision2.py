"""tf32x3 error vs fp64 over the shapes / operand majors of the workload (random-sign operands)."""
import os, sys
sys.path.insert(0, ".")
import torch
from interactron_b200.ops import CudaOps
ops = CudaOps()
torch.manual_seed(0)
print("ITN_GEMM_RZ_COMP =", os.environ.get("ITN_GEMM_RZ_COMP", "(default)"))
def mk(shape, mn):
    if mn:
        return torch.randn(*shape[:-2], shape[-1], shape[-2], device="cuda").transpose(-1, -2)
    return torch.randn(*shape, device="cuda")
cases = [((1805, 256, 2048), 0, 0, ()), ((1805, 2048, 256), 0, 0, ()), ((256, 2048, 1805), 1, 1, (4,)), ((2048, 256, 1805), 1, 1, (4,)),
         ((361, 361, 32), 0, 0, (40,)), ((361, 32, 361), 0, 1, (40,)), ((361, 32, 361), 1, 1, (40,)), ((50, 361, 32), 0, 0, (40,)),
         ((255, 1805, 64), 0, 0, (8,)), ((255, 64, 1805), 0, 1, (8,)), ((1805, 64, 255), 1, 1, (8,)), ((250, 1236, 256), 0, 0, ()),
         ((250, 512, 1496), 0, 0, ()), ((3610, 512, 4608), 0, 0, ()), ((250, 256, 256), 0, 0, (2,)), ((2060, 2060, 64), 0, 0, (8,)),
         ((2060, 64, 2060), 0, 1, (8,)), ((512, 512, 4120), 1, 1, ())]
for (M, N, K), amn, bmn, batch in cases:
    a = mk((*batch, M, K), amn)
    b = mk((*batch, K, N), not bmn)
    ref = a.double() @ b.double()
    out = ops.matmul(a, b)
    t32 = a @ b
    print(f"{M:5d}x{N:5d}x{K:5d} b{batch} A{'MN' if amn else 'K '} B{'MN' if bmn else 'K '}: tf32x3 {((out.double()-ref).norm()/ref.norm()).item():.2e}"
          f" | torch fp32 {((t32.double()-ref).norm()/ref.norm()).item():.2e}")
