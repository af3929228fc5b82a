"""End-to-end CPU check of the drop-in `predict()` orchestration (backbone -> DETR-T -> fusion ->
inner gradient -> clipped SGD step -> re-detect) against the unmodified reference `predict()`,
with the kernels replaced by their torch simulation (oracle/sim_ops.py; the product path has no
such fallback).  fp32 on both sides: agreement is limited by fp32 round-off through the two
different operation orders, hence the 2e-3 / 2e-2 bounds.  Needs /root/reference."""
import pytest
import torch

from oracle import reference_harness as rh
from oracle.sim_ops import SimOps

pytestmark = pytest.mark.skipif(not rh.reference_available(), reason="reference checkout not present")


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("model_type", ["interactron_random", "interactron"])
def test_predict_matches_reference(model_type):
    import interactron_b200 as ib
    from interactron_b200.synthetic import synthetic_episode
    cfg = ib.default_config(model_type, weights="synthetic")
    model = ib.build_model(cfg.MODEL).eval()
    model._ops = SimOps()
    ref = rh.build_reference_model(model_type, model.state_dict())
    data = synthetic_episode(0)
    tr = rh.reference_predict_with_trace(ref, data)
    loop = model._get_loop()
    out = loop.adapt_detect(data["frames"], data["masks"], post_frames=(0,), want_trace=True)
    t = out["trace"]
    assert rel(t["pre_logits"][0], tr["pre"]["pred_logits"][0]) < 1e-4
    assert rel(out["learned_loss"][0], tr["learned_loss"]) < 1e-4
    names = loop.theta_pack.names
    g_ref = torch.cat([g.reshape(-1) for g in tr["grads"]])
    assert rel(t["g"][0], g_ref) < 2e-2
    th_ref = torch.cat([p.reshape(-1) for p in tr["theta_prime"]])
    assert rel(t["theta_prime"][0], th_ref) < 1e-4
    assert len(names) == len(tr["grads"]) == 157
    o = model.predict(data)
    for k in ("pred_logits", "pred_boxes", "box_features", "embedded_memory_features", "image_features"):
        assert o[k].shape == tr["out"][k].shape, k
        assert rel(o[k], tr["out"][k]) < 2e-3, k
    # the module's parameters are untouched by predict()
    for (n1, p1), (n2, p2) in zip(model.state_dict().items(), ref.state_dict().items()):
        assert torch.equal(p1, p2), n1


def test_predict_batches_episodes():
    """b > 1 episodes in one call give the same result as one call per episode."""
    import interactron_b200 as ib
    from interactron_b200.synthetic import collate_episodes, synthetic_episode
    cfg = ib.default_config("interactron_random", weights="synthetic")
    model = ib.build_model(cfg.MODEL).eval()
    model._ops = SimOps()
    eps = [synthetic_episode(i) for i in (1, 2)]
    both = model.predict(collate_episodes(eps))
    for i, e in enumerate(eps):
        one = model.predict(e)
        assert rel(both["pred_logits"][i], one["pred_logits"][0]) < 1e-4
        assert rel(both["pred_boxes"][i], one["pred_boxes"][0]) < 1e-4


def test_policy_steps_match_reference_and_reuse_features():
    """get_next_action over a rollout (1..3 frames seen): equal to the reference's action at every step,
    with the per-frame detector outputs of the previous step reused (host logic of models._policy_logits);
    the batched get_next_actions gives the same actions per episode."""
    import interactron_b200 as ib
    from interactron_b200.synthetic import collate_episodes, synthetic_episode
    cfg = ib.default_config("interactron", weights="synthetic")
    model = ib.build_model(cfg.MODEL).eval()
    model._ops = SimOps()
    ref = rh.build_reference_model("interactron", model.state_dict())
    eps = [synthetic_episode(0), synthetic_episode(4)]

    def cut(d, s):
        o = dict(d)
        o["frames"], o["masks"] = d["frames"][:, :s], d["masks"][:, :s]
        o["category_ids"], o["boxes"] = [d["category_ids"][0][:s]], [d["boxes"][0][:s]]
        return o

    acts = []
    for d in eps:
        h0 = model.policy_cache_hits
        mine = [model.get_next_action(cut(d, s)) for s in (1, 2, 3)]
        assert model.policy_cache_hits == h0 + 2
        with torch.no_grad():
            theirs = [ref.get_next_action(cut(d, s)) for s in (1, 2, 3)]
        assert mine == theirs
        acts.append(mine)
    both = collate_episodes(eps)
    for s in (1, 2, 3):
        o = dict(both)
        o["frames"], o["masks"] = both["frames"][:, :s], both["masks"][:, :s]
        assert model.get_next_actions(o) == [a[s - 1] for a in acts]


def test_policy_action_logits_match_golden_on_the_cpu_simulation():
    """The committed action-logit goldens (tools/make_golden_policy.py, unmodified reference) against the host
    logic run on the CPU simulation of the kernels: pins the fixture itself and `_policy_logits` without a GPU."""
    import os
    import interactron_b200 as ib
    from interactron_b200.synthetic import synthetic_episode
    gold = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "interactron_action_logits.pt"))
    model = ib.build_model(ib.default_config("interactron", weights="synthetic").MODEL).eval()
    model._ops = SimOps()
    data = synthetic_episode(3)
    for s in (1, 2):
        d = dict(data)
        d["frames"], d["masks"] = data["frames"][:, :s], data["masks"][:, :s]
        lg = model._policy_logits(d, 1, s)[0, s - 1].float()
        want = gold["logits"][(3, s)][s - 1]
        assert ((lg - want).norm() / want.norm()).item() < 1e-4
        assert int(lg.argmax()) == gold["actions"][(3, s)]
