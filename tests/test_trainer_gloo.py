"""World-size-2 (gloo, CPU) check of the training iteration's exchange: each rank holds the meta-gradient of
its own episodes; after `parallel.allreduce_meta_grads` (SUM) both ranks run the identical fused trainer step
and end with bit-identical weights equal to a single process that saw the summed gradient."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _build():
    import interactron_b200 as ib
    from oracle.sim_ops import SimOps
    model = ib.build_model(ib.default_config("interactron_random", weights="synthetic").MODEL).eval().double()
    model._ops = SimOps(torch.float64)
    return model


def _grad_for(model, seed):
    loop = model._get_loop()
    sizes = (loop.theta_pack.numel, loop.psi_pack.numel, loop.phi_pack.numel)
    G = torch.zeros(1, sum(sizes), dtype=torch.float64)
    gen = torch.Generator().manual_seed(seed)
    base = 0
    for pack in (loop.theta_pack, loop.psi_pack, loop.phi_pack):
        for nm in pack.names:
            v = pack.view(G[:, base:base + pack.numel], nm)
            v.copy_(torch.randn(v.shape, generator=gen, dtype=torch.float64) * 1e-2)
        base += pack.numel
    return G, sizes


def _step(model, G, sizes):
    from interactron_b200 import meta
    from interactron_b200.trainer import MetaTrainerStep
    tr = MetaTrainerStep(model, 1e-3, 2e-3, 1.0)
    flat = {"all": G, "theta": G[:, :sizes[0]], "psi": G[:, sizes[0]:sizes[0] + sizes[1]], "phi": G[:, sizes[0] + sizes[1]:]}
    model.last_meta_grads = flat
    meta.accumulate_grads(model, flat)
    tr.step()
    return torch.cat([p.detach().reshape(-1) for n, p in model.named_parameters() if not n.startswith("detector.backbone")])


def _worker(rank, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=2)
    torch.set_num_threads(2)
    from interactron_b200 import parallel
    model = _build()
    G, sizes = _grad_for(model, 100 + rank)
    parallel.allreduce_meta_grads(G)
    w = _step(model, G, sizes)
    torch.save(w, os.path.join(out, f"w{rank}.pt"))
    dist.destroy_process_group()


def test_allreduced_trainer_step_world2(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(port, str(tmp_path)), nprocs=2, join=True)
    w0, w1 = torch.load(tmp_path / "w0.pt"), torch.load(tmp_path / "w1.pt")
    assert torch.equal(w0, w1)
    model = _build()
    G0, sizes = _grad_for(model, 100)
    G1, _ = _grad_for(model, 101)
    want = _step(model, G0 + G1, sizes)
    assert torch.equal(w0, want)
