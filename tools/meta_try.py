"""Runs the meta-training step on the GPU a few times and prints losses, timing and launch counts."""
import sys
import time

import torch

sys.path.insert(0, ".")
import interactron_b200 as ib  # noqa: E402
from interactron_b200.synthetic import collate_episodes, synthetic_episode  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "interactron_random"
E = int(sys.argv[2]) if len(sys.argv) > 2 else 2
model = ib.build_model(ib.default_config(name, weights="synthetic").MODEL).cuda().eval()
data = collate_episodes([synthetic_episode(e) for e in range(E)])
ops = model._get_ops()
for it in range(3):
    model.zero_grad(set_to_none=True)
    torch.cuda.synchronize()
    n0, t0 = ops.launch_count(), time.time()
    p, l = model(data, ridx=[(3 * e + 1) % 5 for e in range(E)])
    torch.cuda.synchronize()
    dt = time.time() - t0
    print(f"{name} E={E} iter {it}: {dt * 1e3:.1f} ms, {ops.launch_count() - n0} launches, "
          f"tf32 gemms {ops.n_tf32} simt {ops.n_simt}, peak mem {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB")
print({k: round(float(v), 5) for k, v in l.items()})
gn = {k: float(sum(p.grad.double().pow(2).sum() for n, p in model.named_parameters() if p.grad is not None and n.startswith(k)).sqrt())
      for k in ("detector.", "fusion.")}
print("grad norms", gn)
