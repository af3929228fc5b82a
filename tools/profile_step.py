"""One eager (non-graph) adapt+detect step for ncu: `ncu ... python tools/profile_step.py [E] [workload]`.
Prints the number of launches of the warm-up step so -s/-c can be chosen."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import interactron_b200 as ib  # noqa: E402
from interactron_b200.synthetic import collate_episodes, synthetic_episode  # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 8
name = sys.argv[2] if len(sys.argv) > 2 else "interactron_random"
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
model = ib.build_model(ib.default_config(name, weights="synthetic").MODEL).cuda().eval()
loop = model._get_loop()
d = collate_episodes([synthetic_episode(i, with_targets=False) for i in range(E)])
f, m = d["frames"].cuda(), d["masks"].cuda()

def marker():
    """One launch of a kernel the step never uses: exact step boundaries in the ncu list (tools/_ncu_csv.py)."""
    from interactron_b200 import _lib
    t = torch.zeros(4, device="cuda")
    _lib.check(loop.ops.lib.itn_round_tf32(t.data_ptr(), t.data_ptr(), 4, loop.ops._stream()))


marker()
for s in range(steps):
    n0 = loop.ops.launch_count()
    torch.cuda.nvtx.range_push(f"step{s}")
    loop.adapt_detect(f, m, post_frames=(0,))
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
    print(f"step {s}: {loop.ops.launch_count() - n0} itn launches", flush=True)
    marker()
