"""ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product path.

Runs the *unmodified* allenai/interactron reference (mounted read-only at
/root/reference) on CPU as the parity oracle and as the CPU baseline timer.
Only tests/, tools/make_golden.py, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.

/root/reference exists only in the build container, not on the GPU box: anything
that must run there uses the committed fixtures under tests/golden/ (produced by
tools/make_golden.py with this harness) or oracle/port.py (the CPU restatement).

Harness-side shims (no reference file is edited):
  1. numpy.float alias      — reference models/new_transformer.py:118 uses the removed np.float
  2. is_main_process->False — reference models/detr_models/backbone.py:90 would download weights
  3. torch.load bypass      — reference models/interactron.py:23 loads a checkpoint that is not shipped
Decision D1: `detector.backbone.requires_grad_(False)` before the first call, so
`get_parameters` (reference utils/meta_utils.py:13) yields the same 157 fast weights
the B200 path adapts.  Every parity statement made with this oracle is "D1 mode".
"""
import contextlib
import os
import sys

import torch

REFERENCE_ROOT = os.environ.get("ITN_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "models"))


_loaded = False


def _load():
    global _loaded
    if _loaded:
        return
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    import numpy as np
    if not hasattr(np, "float"):
        np.float = float                                            # shim 1
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import models.detr_models.backbone as bb
    bb.is_main_process = lambda: False                              # shim 2
    _loaded = True


@contextlib.contextmanager
def _no_checkpoint(state_holder):
    """shim 3: `torch.load(config.WEIGHTS)['model']` returns the detector weights we want."""
    real = torch.load
    torch.load = lambda *a, **k: {"model": state_holder["detector"]}
    try:
        yield
    finally:
        torch.load = real


def reference_config(name):
    """Parsed reference config (configs/<name>.yaml) through the reference's own Config class."""
    _load()
    from utils.config_utils import get_config
    return get_config(os.path.join(REFERENCE_ROOT, "configs", name + ".yaml"))


CONFIG_OF = {"interactron": "interactron", "interactron_random": "interactron_random",
             "detr": "single_frame_baseline", "detr_multiframe": "multi_frame_baseline"}


def build_reference_model(model_type, state_dict, freeze_backbone=True):
    """Instantiate the reference model class, load `state_dict` (keys detector.* / fusion.*),
    apply D1, switch to eval()."""
    _load()
    cfg = reference_config(CONFIG_OF[model_type])
    from utils.config_utils import build_model
    det_prefix = "model." if model_type == "detr" else "detector."
    det_sd = {k[len("detector."):]: v for k, v in state_dict.items() if k.startswith("detector.")}
    with _no_checkpoint({"detector": det_sd}):
        model = build_model(cfg.MODEL)
    own = model.state_dict()
    remapped = {}
    for k, v in state_dict.items():
        kk = det_prefix + k[len("detector."):] if k.startswith("detector.") else k
        if kk in own:
            remapped[kk] = v
    missing = [k for k in own if k not in remapped and not k.startswith("criterion.")]
    if missing:
        raise RuntimeError(f"synthetic state_dict lacks reference keys: {missing[:5]} ...")
    model.load_state_dict(remapped, strict=False)
    det = model.model if model_type == "detr" else model.detector
    if freeze_backbone:
        det.backbone.requires_grad_(False)                          # decision D1
    model.eval()
    return model


def reference_fast_weight_names(model):
    """Names of theta in the reference's get_parameters order (by identity matching)."""
    _load()
    from utils.meta_utils import get_parameters
    by_id = {id(p): n for n, p in model.detector.named_parameters()}
    return [by_id[id(p)] for p in get_parameters(model.detector)]


def reference_predict_with_trace(model, data):
    """model.predict(data) while recording the intermediates the parity tests compare:
    pre-adapt detector outputs, learned loss, inner gradient g, adapted weights theta'."""
    _load()
    import utils.meta_utils as mu
    import models.interactron as m_a
    import models.interactron_random as m_b
    trace = {}
    mod = m_a if type(model).__name__ == "interactron" else m_b
    real_sgd = mod.sgd_step

    def spy_sgd(params, grads, lr, clip=0.01):
        out = real_sgd(params, grads, lr, clip)
        trace["grads"] = [g.detach().clone() for g in grads]
        trace["theta_prime"] = [p.detach().clone() for p in out]
        return out

    real_fusion_forward = model.fusion.forward

    def spy_fusion(x):
        trace["pre"] = {k: v.detach().clone() for k, v in x.items()}
        out = real_fusion_forward(x)
        trace["fusion_loss"] = out["loss"].detach().clone()
        trace["fusion_actions"] = out["actions"].detach().clone()
        return out

    mod.sgd_step = spy_sgd
    model.fusion.forward = spy_fusion
    try:
        out = model.predict(data)
    finally:
        mod.sgd_step = real_sgd
        model.fusion.forward = real_fusion_forward
    trace["learned_loss"] = torch.norm(trace["fusion_loss"])
    trace["out"] = {k: v.detach().clone() for k, v in out.items()}
    del mu
    return trace


def reference_forward_with_grads(model, data, ridx):
    """model(data) (the meta-training step, reference models/interactron.py:61-151) with the
    `random.randint(0, 4)` draws replaced by `ridx` (one per episode).  Returns (predictions, losses,
    {parameter name: grad or None}); the model's .grad fields are cleared afterwards."""
    _load()
    import random
    draws = list(ridx)
    real = random.randint
    random.randint = lambda a, b: draws.pop(0)
    try:
        model.zero_grad(set_to_none=True)
        preds, losses = model(data)
    finally:
        random.randint = real
    grads = {n: (None if p.grad is None else p.grad.detach().clone()) for n, p in model.named_parameters()}
    model.zero_grad(set_to_none=True)
    return ({k: v.detach().clone() for k, v in preds.items()}, {k: v.detach().clone() for k, v in losses.items()},
            grads)
