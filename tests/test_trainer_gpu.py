"""GPU parity of the trainer step (csrc/itn_trainer.cu behind trainer.MetaTrainerStep): the fused
global-norm clip + two Adam optimisers + zero_grad against the reference's own sequence
(engine/interactron_trainer.py:70-71,106-110) executed by torch on the same device, and against its
float64 restatement (oracle/sim_ops.py).  fp32 tolerance: 2e-6 relative on moments and norm, 1e-5 of
the applied update on the weights."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("n", [4, 1003, 1 << 20, 14_798_296 + 7])
def test_clip_adam_kernel_vs_float64(n):
    from interactron_b200.ops import CudaOps
    from interactron_b200.trainer import _mark_no_grad
    from oracle.sim_ops import SimOps
    ops, sim = CudaOps(), SimOps(torch.float64)
    gen = torch.Generator(device="cuda").manual_seed(n)
    w = torch.randn(n, generator=gen, device="cuda")
    m = torch.randn(n, generator=gen, device="cuda") * 1e-2
    v = torch.rand(n, generator=gen, device="cuda") * 1e-3
    for step, scale, max_norm in ((1, 5.0, 1.0), (7, 1e-4, 1.0), (300, 1.0, 0.0)):
        g = torch.randn(n, generator=gen, device="cuda") * scale
        if n > 8:
            _mark_no_grad(g[4:8])                       # "no gradient" slots are skipped everywhere
        w64, g64, m64, v64 = (t.double().cpu() for t in (w, g, m, v))
        part = ops.sumsq_partials(g)
        norm = torch.zeros(1, device="cuda")
        ops.clip_adam_step_(w, g, m, v, part, max_norm, 3e-3, (0.9, 0.999), 1e-8, step, zero_grad=(step == 7), norm_out=norm)
        n64 = torch.zeros(1, dtype=torch.float64)
        sim.clip_adam_step_(w64, g64, m64, v64, sim.sumsq_partials(g64), max_norm, 3e-3, (0.9, 0.999), 1e-8, step,
                            zero_grad=(step == 7), norm_out=n64)
        torch.cuda.synchronize()
        assert float(norm) == pytest.approx(float(n64), rel=2e-6)
        assert rel(m, m64) < 2e-6 and rel(v, v64) < 2e-6 and rel(w, w64) < 2e-6
        if step == 7:
            assert float(g.abs().sum()) == 0.0
        if n > 8:
            assert torch.equal(w[4:8].cpu().double(), w64[4:8])


@pytest.mark.parametrize("model_type", ["interactron_random", "interactron"])
def test_trainer_iteration_matches_torch(model_type):
    """forward() -> MetaTrainerStep.step() twice vs clip_grad_norm_ + 2x Adam of torch on a copy; then
    predict() with the updated weights equals a fresh model loaded from the torch-updated state_dict."""
    import interactron_b200 as ib
    from interactron_b200.synthetic import collate_episodes, synthetic_episode
    from interactron_b200.trainer import MetaTrainerStep
    cfg = ib.default_config(model_type, weights="synthetic")
    model = ib.build_model(cfg.MODEL).cuda().eval()
    ref = copy.deepcopy(model)
    lr_d, lr_s, clip = 1e-4, 1e-3, 1.0          # larger than the YAML's 1e-5 / 1e-4 so the update is well above fp32 noise
    opt_d = torch.optim.Adam(ref.detector.parameters(), lr=lr_d)
    opt_s = torch.optim.Adam(ref.fusion.parameters(), lr=lr_s)
    tr = MetaTrainerStep(model, lr_d, lr_s, clip)
    probe = collate_episodes([synthetic_episode(40, with_targets=False)])
    out0 = {k: v.clone() for k, v in model.predict(probe).items()}
    for it in range(2):
        data = collate_episodes([synthetic_episode(20 + it)])
        before = {n: p.detach().clone() for n, p in model.named_parameters()}
        model(data, ridx=[it])
        grads = {n: (None if p.grad is None else p.grad.detach().clone()) for n, p in model.named_parameters()}
        for n, p in ref.named_parameters():
            p.grad = grads[n]
        want_norm = torch.nn.utils.clip_grad_norm_(ref.parameters(), clip)
        opt_d.step(); opt_s.step(); opt_d.zero_grad(); opt_s.zero_grad()
        got = tr.step(n_frames=5)
        torch.cuda.synchronize()
        assert float(got["grad_norm"]) == pytest.approx(float(want_norm), rel=1e-5)
        assert all(p.grad is None for p in model.parameters())
        num = den = 0.0
        for (n, p), (_, q) in zip(model.named_parameters(), ref.named_parameters()):
            if grads[n] is None:
                assert torch.equal(p, before[n]), n
                continue
            num += float((p.detach().double() - q.detach().double()).pow(2).sum())
            den += float((q.detach().double() - before[n].double()).pow(2).sum())
        assert den > 0 and (num / den) ** 0.5 < 1e-5, (it, (num / den) ** 0.5)
    # the updated weights are what the hot path now runs on (flat buffers + W^T twins + CUDA graphs)
    out1 = model.predict(probe)
    assert rel(out1["pred_logits"], out0["pred_logits"]) > 1e-6
    fresh = ib.build_model(cfg.MODEL).cuda().eval()
    fresh.load_state_dict(ref.state_dict())
    out2 = fresh.predict(probe)
    for k in ("pred_logits", "pred_boxes"):
        assert rel(out1[k], out2[k]) < 2e-4, k


def test_checkpoint_accumulate_bit_exact_and_windowed_average():
    """itn_ckpt_accumulate == torch's `saved = w * v` / `saved += w * v` bit for bit (no FMA contraction), and the
    averager's state_dict equals the reference arithmetic (engine/interactron_trainer.py:48-57) on real snapshots."""
    import interactron_b200 as ib
    from interactron_b200.ops import CudaOps
    from interactron_b200.trainer import CheckpointAverager
    ops = CudaOps()
    g = torch.Generator(device="cuda").manual_seed(0)
    x1 = torch.randn(1_000_003, device="cuda", generator=g)
    x2 = torch.randn(1_000_003, device="cuda", generator=g)
    acc = torch.full_like(x1, float("nan"))
    w = 1.0 / 3.0
    ops.ckpt_accumulate_(acc, x1, w, True)
    want = torch.tensor(w, dtype=torch.float32, device="cuda") * x1
    assert torch.equal(acc, want)
    ops.ckpt_accumulate_(acc, x2, w, False)
    want += torch.tensor(w, dtype=torch.float32, device="cuda") * x2
    assert torch.equal(acc, want)

    model = ib.build_model(ib.default_config("interactron_random", weights="synthetic").MODEL).cuda().eval()
    avg = CheckpointAverager(model)
    saved, window = None, 2
    for epoch in range(window):
        with torch.no_grad():
            for p in model.parameters():
                if p.requires_grad:
                    p.add_(0.01 * (epoch + 1))
        sd = {k: v.clone() for k, v in model.state_dict().items()}
        saved = {k: (1.0 / window) * v for k, v in sd.items()} if saved is None else \
            {k: saved[k] + (1.0 / window) * v for k, v in sd.items()}
        avg.record_checkpoint(1.0 / window)
    out = avg.state_dict()
    assert list(out.keys()) == list(saved.keys())
    for k in saved:
        assert torch.equal(out[k], saved[k]), k
