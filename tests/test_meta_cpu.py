"""End-to-end CPU check (float64) of the meta-training step `forward(data)` - pre-adapt pass, inner
gradient, clipped SGD step, post-adapt passes, matcher + criterion, second-order supervisor
gradients (dual-number pass) and first-order detector gradients - against the unmodified
reference's `forward()` and the `.grad` it leaves on every Parameter (D1 mode, eval()).
Kernels are replaced by their torch simulation (oracle/sim_ops.py).  Needs /root/reference."""
import pytest
import torch

from oracle import reference_harness as rh
from oracle.sim_ops import SimOps

pytestmark = pytest.mark.skipif(not rh.reference_available(), reason="reference checkout not present")


def _to_double(data):
    out = dict(data)
    out["frames"] = data["frames"].double()
    out["boxes"] = [[b.double() for b in ep] for ep in data["boxes"]]
    return out


@pytest.mark.parametrize("model_type,eps", [("interactron_random", (0, 1)), ("interactron", (0,))])
def test_forward_matches_reference(model_type, eps):
    import interactron_b200 as ib
    from interactron_b200.synthetic import collate_episodes, synthetic_episode
    cfg = ib.default_config(model_type, weights="synthetic")
    model = ib.build_model(cfg.MODEL).eval()
    ref = rh.build_reference_model(model_type, model.state_dict()).double()
    model = model.double()
    sim = SimOps(torch.float64)
    model._ops = sim
    model.criterion._ops_override = sim
    model.criterion.matcher._ops_override = sim
    data = _to_double(collate_episodes([synthetic_episode(e) for e in eps]))
    ridx = [3, 1][:len(eps)]
    rounds = 2 if model_type == "interactron" else 1      # second round: path storage already populated
    for _ in range(rounds):
        torch.set_default_dtype(torch.float64)       # fusion B allocates with the default dtype
        try:
            p_ref, l_ref, g_ref = rh.reference_forward_with_grads(ref, data, ridx)
        finally:
            torch.set_default_dtype(torch.float32)
        model.zero_grad(set_to_none=True)
        p, l = model(data, ridx=ridx)
        assert list(l.keys()) == list(l_ref.keys())
        for k in l_ref:
            assert float(l[k]) == pytest.approx(float(l_ref[k]), rel=1e-6, abs=1e-9), k
        for k in ("pred_logits", "pred_boxes"):
            assert p[k].shape == p_ref[k].shape
            assert float((p[k] - p_ref[k]).norm() / p_ref[k].norm()) < 1e-8, k
        bad, n = [], 0
        for name, prm in model.named_parameters():
            gr = g_ref[name]
            if gr is None:
                assert prm.grad is None, name
                continue
            assert prm.grad is not None, name
            n += 1
            err = float((prm.grad - gr).norm())
            if err > 1e-7 * float(gr.norm()) + 1e-12:
                bad.append((name, err / max(float(gr.norm()), 1e-30)))
        assert not bad, sorted(bad, key=lambda t: -t[1])[:10]
        assert n > 250
