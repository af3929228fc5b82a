"""GPU parity of the drop-in models against the reference goldens (tests/golden, produced by the
unmodified reference on CPU, D1 mode) and against the travelling CPU oracle (oracle/port.py).

Bar (BASELINE.json north_star): logits, boxes and adapted weights within 1e-3 relative L2;
matcher assignments bit-exact.  Default precision: tf32x3 GEMMs, fp32 cuDNN backbone.
"""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-3


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def build(name):
    import interactron_b200 as ib
    return ib.build_model(ib.default_config(name, weights="synthetic").MODEL).cuda().eval()


@pytest.fixture(scope="module")
def rand_model():
    return build("interactron_random")


@pytest.fixture(scope="module")
def full_model():
    return build("interactron")


@pytest.mark.parametrize("name", ["interactron_random", "interactron"])
def test_adapt_detect_trace_matches_reference(name, rand_model, full_model):
    from interactron_b200.synthetic import synthetic_episode
    model = rand_model if name == "interactron_random" else full_model
    gold = torch.load(os.path.join(GOLD, f"{name}_predict.pt"))
    loop = model._get_loop()
    assert loop.ops.precision == "tf32x3" and not loop.backbone_tf32
    for ep, g in gold["episodes"].items():
        d = synthetic_episode(ep)
        out = loop.adapt_detect(d["frames"].cuda(), d["masks"].cuda(), post_frames=(0,), want_trace=True)
        t = out["trace"]
        assert rel(t["pre_logits"][0, 0], g["pre_logits_f0"]) < TOL
        assert rel(t["pre_logits"][0, 4], g["pre_logits_f4"]) < TOL
        assert rel(t["pre_boxes"][0], g["pre_boxes"]) < TOL
        assert rel(t["loss_vec"][0], g["loss_vec"]) < TOL
        assert rel(out["learned_loss"][0], g["learned_loss"]) < TOL
        assert rel(out["actions"][0], g["actions"]) < TOL
        names = gold["theta_names"]
        assert names == loop.theta_pack.names
        gn = torch.stack([loop.theta_pack.view(t["g"], n)[0].norm() for n in names]).cpu()
        assert ((gn - g["g_norms"]).abs() / g["g_norms"]).max().item() < 5e-3
        # adapted weights: every stored tensor within 1e-3, and all 157 norms
        for n, (gr, tp) in g["small"].items():
            assert rel(loop.theta_pack.view(t["theta_prime"], n)[0], tp) < TOL, n
        tn = torch.stack([loop.theta_pack.view(t["theta_prime"], n)[0].double().norm() for n in names]).cpu()
        assert ((tn - g["theta_prime_norms"]).abs() / g["theta_prime_norms"]).max().item() < 1e-5
        assert rel(out["pred_logits"][0], g["pred_logits"][0]) < TOL
        assert rel(out["pred_boxes"][0], g["pred_boxes"][0]) < TOL
        assert rel(out["box_features"][0], g["box_features"][0]) < TOL


@pytest.mark.parametrize("name", ["interactron_random", "interactron"])
def test_padded_frames_match_reference(name, rand_model, full_model):
    """Non-zero masks (padded frames): key-padding masks in every encoder / cross attention and the
    mask-dependent sine position embedding, against the reference golden (tools/make_golden_masked.py)."""
    from interactron_b200.synthetic import collate_episodes, masked_episode, synthetic_episode
    rand_model = rand_model if name == "interactron_random" else full_model
    gold = torch.load(os.path.join(GOLD, f"{name}_predict_masked.pt"))
    loop = rand_model._get_loop()
    for ep, g in gold.items():
        d = masked_episode(ep)
        assert int(d["masks"].sum()) > 0
        out = loop.adapt_detect(d["frames"].cuda(), d["masks"].cuda(), post_frames=(0,), want_trace=True)
        t = out["trace"]
        assert rel(t["pre_logits"][0], g["pre_logits"]) < TOL
        assert rel(t["pre_boxes"][0], g["pre_boxes"]) < TOL
        assert rel(out["learned_loss"][0], g["learned_loss"]) < TOL
        gn = torch.stack([loop.theta_pack.view(t["g"], n)[0].norm() for n in loop.theta_pack.names]).cpu()
        assert ((gn - g["g_norms"]).abs() / g["g_norms"]).max().item() < 5e-3
        assert rel(out["pred_logits"][0], g["pred_logits"][0]) < TOL
        assert rel(out["pred_boxes"][0], g["pred_boxes"][0]) < TOL
        # the mask must matter (an unmasked run of the same pixels differs), and a padded episode in a
        # batch with an unpadded one gives the same answer as alone (per-frame masks, ragged batch)
        d0 = dict(d)
        d0["masks"] = torch.zeros_like(d["masks"])
        o0 = rand_model.predict(d0)
        assert rel(o0["pred_logits"][0], g["pred_logits"][0]) > 10 * TOL
        both = rand_model.predict(collate_episodes([synthetic_episode(1), d]))
        assert rel(both["pred_logits"][1], g["pred_logits"][0]) < TOL
        assert rel(both["pred_boxes"][1], g["pred_boxes"][0]) < TOL


def test_predict_public_api_graph_and_eager(rand_model):
    """predict(data) with host tensors: CUDA-graph replay == eager launches, shapes as the reference."""
    from interactron_b200.synthetic import synthetic_episode
    gold = torch.load(os.path.join(GOLD, "interactron_random_predict.pt"))["episodes"]
    m = rand_model
    for use_graph in (False, True, True):
        m.use_cuda_graph = use_graph
        for ep in (0, 1):
            o = m.predict(synthetic_episode(ep))
            assert o["pred_logits"].shape == (1, 1, 50, 1236) and o["pred_boxes"].shape == (1, 1, 50, 4)
            assert o["image_features"].shape == (1, 1, 2048, 19, 19)
            assert o["embedded_memory_features"].shape == (1, 1, 256, 19, 19)
            assert o["box_features"].shape == (1, 1, 50, 256)
            assert rel(o["pred_logits"], gold[ep]["pred_logits"]) < TOL
            assert rel(o["pred_boxes"], gold[ep]["pred_boxes"]) < TOL
    assert any(k[0] == "predict" for k in m._graphs)


def test_predict_batch_equals_single_and_is_deterministic(rand_model):
    from interactron_b200.synthetic import collate_episodes, synthetic_episode
    m = rand_model
    eps = [synthetic_episode(i) for i in (0, 1, 2, 3)]
    both = m.predict(collate_episodes(eps))
    again = m.predict(collate_episodes(eps))
    assert torch.equal(both["pred_logits"], again["pred_logits"])        # bit-identical replays
    for i, e in enumerate(eps):
        one = m.predict(e)
        assert rel(both["pred_logits"][i], one["pred_logits"][0]) < 1e-4
        assert rel(both["pred_boxes"][i], one["pred_boxes"][0]) < 1e-4


def test_predict_at_the_bench_batch_size_matches_reference_and_singles(rand_model):
    """The bench's full configuration (62 episodes per step, BASELINE configs[2]): episodes 0 and 1 of the batch
    equal the reference goldens, sampled episodes equal their single-episode predict() (independence of the
    episodes inside the batched kernels: per-episode fast weights, tile tails, CTA-pair and batched-layer GEMMs
    all take their large-batch paths here), and a replay is bit-identical."""
    from interactron_b200.synthetic import collate_episodes, synthetic_episode
    gold = torch.load(os.path.join(GOLD, "interactron_random_predict.pt"))["episodes"]
    m = rand_model
    E = 62
    batch = collate_episodes([synthetic_episode(i, with_targets=False) for i in range(E)])
    out = m.predict(batch)
    assert out["pred_logits"].shape == (E, 1, 50, 1236)
    for ep in (0, 1):
        assert rel(out["pred_logits"][ep], gold[ep]["pred_logits"][0]) < TOL
        assert rel(out["pred_boxes"][ep], gold[ep]["pred_boxes"][0]) < TOL
    for ep in (5, 37, 61):
        one = m.predict(synthetic_episode(ep, with_targets=False))
        assert rel(out["pred_logits"][ep], one["pred_logits"][0]) < 1e-4
        assert rel(out["pred_boxes"][ep], one["pred_boxes"][0]) < 1e-4
    again = m.predict(batch)
    assert torch.equal(out["pred_logits"], again["pred_logits"]) and torch.equal(out["pred_boxes"], again["pred_boxes"])
    m._graphs.clear()
    torch.cuda.empty_cache()


def test_pipelined_host_input_is_bit_identical(rand_model):
    """predict() on host frames hides the H2D copy behind the trunk (graph.PipelinedPredict: trunk on the first
    quarter of the frames, trunk on the rest, then the step, as three graphs fed by a copy stream).  Same bits as
    the copy-first single-graph path, for pageable and pinned inputs, over repeated calls with different data."""
    from interactron_b200.synthetic import collate_episodes, synthetic_episode
    m = rand_model
    m.pipelined_input = True          # opt-in (off by default: no gain on the power-capped board)
    batches = [collate_episodes([synthetic_episode(i, with_targets=False) for i in range(k, k + 16)]) for k in (0, 20)]
    piped = []
    for b in batches + batches[:1]:
        pinned = dict(b)
        pinned["frames"] = b["frames"].pin_memory()
        o = m.predict(pinned)
        piped.append({k: v.clone() for k, v in o.items()})
    assert any(k[0] == "predict_pipe" for k in m._graphs)
    assert all(torch.equal(piped[0][k], piped[2][k]) for k in piped[0])
    o_pageable = m.predict(batches[1])
    assert all(torch.equal(o_pageable[k], piped[1][k]) for k in o_pageable)
    m.pipelined_input = False
    try:
        for b, want in zip(batches, piped):
            o = m.predict(b)
            for k in want:
                assert torch.equal(o[k], want[k]), k
    finally:
        m.pipelined_input = False
    m._graphs.clear()
    torch.cuda.empty_cache()


def test_predict_leaves_parameters_intact_and_tracks_updates(rand_model):
    from interactron_b200.synthetic import synthetic_episode
    m = rand_model
    before = {k: v.clone() for k, v in m.state_dict().items()}
    d = synthetic_episode(0)
    o1 = m.predict(d)
    for k, v in m.state_dict().items():
        assert torch.equal(v, before[k]), k
    # an in-place optimizer-style update must be picked up (flat buffers are re-packed)
    p = m.detector.class_embed.bias
    with torch.no_grad():
        p.add_(0.5)
    o2 = m.predict(d)
    assert (o2["pred_logits"] - o1["pred_logits"]).abs().mean().item() > 0.1
    with torch.no_grad():
        p.sub_(0.5)
    o3 = m.predict(d)
    # (b + 0.5) - 0.5 != b in fp32 (off by up to 3e-8), and the adaptation amplifies it: not bit-equal
    assert rel(o3["pred_logits"], o1["pred_logits"]) < 5e-4


def test_live_cross_check_against_cpu_oracle(rand_model):
    """Episode that is NOT in the goldens: CUDA path vs oracle/port.py run on the host."""
    from oracle import port
    from interactron_b200.synthetic import synthetic_episode
    m = rand_model
    d = synthetic_episode(11)
    o = m.predict(d)
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    import copy
    body = copy.deepcopy(m.detector.backbone[0].body).cpu()
    ref = port.predict(sd, body, d, "B", lr=m.config.ADAPTIVE_LR)
    assert rel(o["pred_logits"], ref["pred_logits"]) < TOL
    assert rel(o["pred_boxes"], ref["pred_boxes"]) < TOL


def test_policy_actions_match_reference(full_model):
    from interactron_b200.synthetic import synthetic_episode
    acts = torch.load(os.path.join(GOLD, "interactron_actions.pt"))
    data = synthetic_episode(0)
    for s in range(1, 5):
        d = dict(data)
        d["frames"], d["masks"] = data["frames"][:, :s], data["masks"][:, :s]
        assert full_model.get_next_action(d) == acts[s]


def test_policy_action_logits_match_reference(full_model):
    """A1 pinned on the continuous quantity: the action LOGITS of the policy step (reference
    models/interactron.py:174-197, `fusion_out['actions']`) for 4 episodes x 1..4 frames seen, against the
    unmodified reference (tests/golden/interactron_action_logits.pt, tools/make_golden_policy.py) to 1e-3
    relative - with random-init weights the argmax is the same action everywhere, the logits are not
    (they differ by ~1 % between episodes) - through both the single-episode and the lock-step batched call."""
    from interactron_b200.synthetic import collate_episodes, synthetic_episode
    gold = torch.load(os.path.join(GOLD, "interactron_action_logits.pt"))
    eps = {ep: synthetic_episode(ep) for ep in gold["episodes"]}

    def cut(d, s):
        o = dict(d)
        o["frames"], o["masks"] = d["frames"][:, :s], d["masks"][:, :s]
        return o

    worst = 0.0
    for ep, data in eps.items():
        for s in range(1, 5):
            lg = full_model._policy_logits(cut(data, s), 1, s)[0, s - 1].float().cpu()
            want = gold["logits"][(ep, s)][s - 1]
            worst = max(worst, rel(lg, want))
            assert rel(lg, want) < TOL, (ep, s, lg, want)
            assert int(lg.argmax()) == gold["actions"][(ep, s)]
            assert full_model.get_next_action(cut(data, s)) == gold["actions"][(ep, s)]
    batch = collate_episodes([eps[ep] for ep in gold["episodes"]])
    for s in (2, 4):
        lg = full_model._policy_logits(cut(batch, s), len(eps), s)[:, s - 1].float().cpu()
        for i, ep in enumerate(gold["episodes"]):
            assert rel(lg[i], gold["logits"][(ep, s)][s - 1]) < TOL
    # the goldens discriminate: neighbouring episodes are further apart than the tolerance
    a, b = gold["logits"][(0, 1)][0], gold["logits"][(3, 1)][0]
    assert rel(a, b) > 5 * TOL and worst < TOL


def test_batched_policy_step_equals_single_episodes(full_model):
    """get_next_actions (extension: b environments in lock-step) == get_next_action per episode."""
    from interactron_b200.synthetic import collate_episodes, synthetic_episode
    eps = [synthetic_episode(e) for e in (0, 5, 6)]
    for s in (1, 3):
        def cut(d):
            o = dict(d)
            o["frames"], o["masks"] = d["frames"][:, :s], d["masks"][:, :s]
            return o
        singles = [full_model.get_next_action(cut(d)) for d in eps]
        batch = cut(collate_episodes(eps))
        assert full_model.get_next_actions(batch) == singles
        one = torch.cat([full_model._policy_logits(cut(d), 1, s).clone() for d in eps])
        assert rel(full_model._policy_logits(batch, 3, s), one) < 1e-4


def test_policy_rollout_reuses_per_frame_features(full_model):
    """The policy steps of a rollout (s = 1..4 frames of the same episodes) run the detector only on the new
    frame; same action logits as recomputing every frame; a changed prefix or changed weights miss."""
    from interactron_b200.synthetic import collate_episodes, synthetic_episode
    m = full_model
    batch = collate_episodes([synthetic_episode(e) for e in (7, 8)])

    def cut(d, s):
        o = dict(d)
        o["frames"], o["masks"] = d["frames"][:, :s], d["masks"][:, :s]
        return o

    m.policy_cache = False
    want = [m._policy_logits(cut(batch, s), 2, s).clone() for s in range(1, 5)]
    m.policy_cache, m._policy_cache = True, None
    h0 = m.policy_cache_hits
    got = [m._policy_logits(cut(batch, s), 2, s).clone() for s in range(1, 5)]
    assert m.policy_cache_hits == h0 + 3
    for a, b in zip(got, want):
        assert rel(a, b) < 1e-5
    # another episode's frames with the same geometry: prefix differs -> full recompute, correct answer
    other = collate_episodes([synthetic_episode(e) for e in (9, 8)])
    m._policy_logits(cut(batch, 2), 2, 2)
    h1 = m.policy_cache_hits
    o3 = m._policy_logits(cut(other, 3), 2, 3).clone()
    assert m.policy_cache_hits == h1
    m.policy_cache = False
    assert rel(o3, m._policy_logits(cut(other, 3), 2, 3)) < 1e-5
    m.policy_cache = True
    # a weight update between two steps invalidates the cache
    m._policy_logits(cut(batch, 1), 2, 1)
    with torch.no_grad():
        m.detector.class_embed.bias.add_(0.01)
    h2 = m.policy_cache_hits
    m._policy_logits(cut(batch, 2), 2, 2)
    assert m.policy_cache_hits == h2
    with torch.no_grad():
        m.detector.class_embed.bias.sub_(0.01)


def test_baselines_match_reference():
    from interactron_b200.synthetic import synthetic_episode
    base = torch.load(os.path.join(GOLD, "baselines_predict.pt"))
    o = build("single_frame_baseline").predict(synthetic_episode(0, frames=1))
    for k, v in base["detr_ep0_1frame"].items():
        assert o[k].shape == v.shape and rel(o[k], v) < TOL, k
    o = build("multi_frame_baseline").predict(synthetic_episode(0))
    for k, v in base["detr_multiframe_ep0"].items():
        assert o[k].shape == v.shape and rel(o[k], v) < TOL, k


def test_matcher_assignments_bit_exact():
    """Cost matrix from the CUDA kernel -> scipy LSAP == the reference HungarianMatcher's indices."""
    from scipy.optimize import linear_sum_assignment
    from interactron_b200.ops import CudaOps
    ops = CudaOps()
    gold = torch.load(os.path.join(GOLD, "matcher_assignments.pt"))
    for seed, ref_idx in gold.items():
        gen = torch.Generator().manual_seed(100 + seed)
        logits = torch.randn(5, 50, 1236, generator=gen)
        boxes = torch.rand(5, 50, 4, generator=gen) * 0.5 + 0.1
        labels, tboxes, off = [], [], [0]
        for f in range(5):
            n = int(torch.randint(3, 9, (1,), generator=gen))
            labels.append(torch.randint(1, 1235, (n,), generator=gen))
            tboxes.append(torch.cat([torch.rand(n, 2, generator=gen) * 0.6 + 0.2,
                                     torch.rand(n, 2, generator=gen) * 0.3 + 0.05], 1))
            off.append(off[-1] + n)
        cost = ops.matcher_cost(logits.cuda(), boxes.cuda(), torch.cat(tboxes).cuda().contiguous(),
                                torch.cat(labels).cuda(), torch.tensor(off, dtype=torch.int32).cuda(),
                                1.0, 5.0, 2.0).cpu()
        for f in range(5):
            n = off[f + 1] - off[f]
            c = cost[50 * off[f]:50 * off[f + 1]].view(50, n)
            i, j = linear_sum_assignment(c.numpy())
            assert torch.equal(torch.as_tensor(i), ref_idx[f][0]) and torch.equal(torch.as_tensor(j), ref_idx[f][1])


@pytest.mark.xfail(strict=True, reason="decision D1: this repo freezes the backbone (157 fast weights); the UNMODIFIED "
                                       "reference also adapts ResNet layer2-4 (199 tensors, reference "
                                       "models/detr_models/backbone.py:61-63) - SURVEY.md section 8f-3, not built")
def test_unmodified_reference_without_d1(rand_model):
    """Yardstick of the D1 gap: predict() against the reference run WITHOUT freezing the backbone
    (tools/make_golden_unfrozen.py).  On the synthetic weights D1 itself moves the logits by 13 % and the boxes
    by 3 % (stored in the fixture), far outside the 1e-3 bar - every parity statement of this repo is D1 mode."""
    from interactron_b200.synthetic import synthetic_episode
    gold = torch.load(os.path.join(GOLD, "interactron_random_predict_unfrozen.pt"))
    assert gold["n_theta"] == 199 and gold["n_theta_backbone"] == 42
    out = rand_model.predict({k: (v.cuda() if torch.is_tensor(v) else v)
                              for k, v in synthetic_episode(gold["episode"]).items()})
    assert rel(out["pred_logits"], gold["pred_logits"]) < TOL
    assert rel(out["pred_boxes"], gold["pred_boxes"]) < TOL
