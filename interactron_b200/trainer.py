"""The optimiser half of the reference's meta-training iteration, on the flat buffers.

Reference (engine/interactron_trainer.py): two `torch.optim.Adam` (detector parameters at
DETECTOR_LR, fusion parameters at SUPERVISOR_LR, :70-71), and per iteration (:106-124)

    torch.nn.utils.clip_grad_norm_(model.parameters(), GRAD_NORM_CLIP)
    detector_optimizer.step(); supervisor_optimizer.step()
    detector_optimizer.zero_grad(); supervisor_optimizer.zero_grad()
    [LR_DECAY: linear warm-up / cosine decay of the supervisor learning rate by frames seen]

= ~330 parameters x (norm, scale, 2 moment updates, addcdiv).  `forward()` already leaves the
whole meta-gradient in ONE flat buffer [theta | psi | phi] (meta.py) and the weights live in three
flat buffers (episode.InnerLoop), so the iteration is 4 launches: `itn_sumsq_partials` over the
gradient buffer and one `itn_clip_adam_step` per weight buffer (global-norm clip coefficient,
both Adam moments, the weight update and zero_grad fused; csrc/itn_trainer.cu).  The model's
Parameters are re-pointed at views of the flat weight buffers, so `state_dict()`, checkpoints and
the next `forward()` / `predict()` see the update without a scatter; only the W^T twins of the
data-gradient GEMMs are rebuilt.
"""
import math

import torch

NO_GRAD_BITS = 0x7FC0DEAD      # ITN_NO_GRAD_BITS (include/interactron_b200.h)


def _mark_no_grad(t):
    """Fill contiguous `t` with the "this parameter has no gradient" marker.  Written as integer bits:
    a float fill could canonicalise the NaN payload on its way to the device."""
    if t.dtype == torch.float32:
        t.view(torch.int32).fill_(NO_GRAD_BITS)
    else:                                   # float64 simulator backend: any NaN is the marker
        t.fill_(float("nan"))


class MetaTrainerStep:
    def __init__(self, model, detector_lr, supervisor_lr, grad_norm_clip, betas=(0.9, 0.999), eps=1e-8,
                 lr_decay=False, warmup_tokens=0, final_tokens=1):
        self.model = model
        self.detector_lr, self.base_supervisor_lr = float(detector_lr), float(supervisor_lr)
        self.supervisor_lr = float(supervisor_lr)
        self.max_norm, self.betas, self.eps = float(grad_norm_clip), (float(betas[0]), float(betas[1])), float(eps)
        self.lr_decay, self.warmup_tokens, self.final_tokens = bool(lr_decay), float(warmup_tokens), float(final_tokens)
        self.tokens = 0
        self.t = 0
        self._slot_steps = None             # per-parameter Adam step counts (host ints), see _flat_grads
        self.loop = model._get_loop()
        self._own_g = None
        self._alias_parameters()
        n = self._sizes()
        ops = self.loop.ops
        self.m = ops.zeros(sum(n))
        self.v = ops.zeros(sum(n))
        self.norm = ops.zeros(1)

    @classmethod
    def from_config(cls, model, trainer_cfg):
        """trainer_cfg = cfg.TRAINER of configs/interactron*.yaml (BETA1/BETA2/WEIGHT_DECAY there are
        not used by the reference trainer either: it builds Adam with PyTorch's defaults)."""
        g = lambda k, d: getattr(trainer_cfg, k, d)
        return cls(model, trainer_cfg.DETECTOR_LR, trainer_cfg.SUPERVISOR_LR, trainer_cfg.GRAD_NORM_CLIP,
                   lr_decay=g("LR_DECAY", False), warmup_tokens=g("WARMUP_TOKENS", 0), final_tokens=g("FINAL_TOKENS", 1))

    # ------------------------------------------------------------------ layout
    def _segments(self):
        lp = self.loop
        return (("theta", lp.theta_pack, lp.theta_params, lp.theta), ("psi", lp.psi_pack, lp.psi_params, lp.psi),
                ("phi", lp.phi_pack, lp.phi_params, lp.phi))

    def _sizes(self):
        return [pack.numel for _, pack, _, _ in self._segments()]

    def _alias_parameters(self):
        """Parameter.data <- view of its slot in the flat buffer (values are identical: the buffers were
        packed from these Parameters).  Caches (parameter, data pointer, byte offset of its gradient slot)."""
        model, lp = self.model, self.loop
        if model._loop_key != model._weights_key():
            lp.refresh_weights()
        self._slots = []
        base = 0
        for _, pack, params, flat in self._segments():
            for name, p in zip(pack.names, params):
                view = pack.view(flat, name)[0]
                if p.data.data_ptr() != view.data_ptr():
                    p.data = view
                self._slots.append((p, view.data_ptr(), (base + pack.offsets[name]) * flat.element_size(), base, pack, name))
            base += pack.numel
        self._all_params = list(model.parameters())
        self._twin_graph = None
        model._loop_key = model._weights_key()

    def _aliased(self):
        return all(s[0].data_ptr() == s[1] for s in self._slots)

    def _flat_grads(self):
        """The flat gradient buffer [theta | psi | phi] behind the Parameters' .grad.  After one
        `forward()` since the last zero_grad every .grad is a view of `model.last_meta_grads["all"]`
        (zero copies); anything else (several accumulated forwards into foreign tensors, grads set by
        hand) is gathered into an own buffer.  Parameters without a gradient get the no-gradient marker."""
        # torch.optim.Adam keeps one step count PER parameter and advances it only when .grad is not None; the
        # fused step applies ONE count (self.t + 1) to every parameter it updates.  The two agree as long as every
        # parameter updated now was also updated on all earlier steps - tracked per parameter and checked, not
        # assumed (a parameter that stops receiving gradients is fine until it receives one again).
        if self._slot_steps is None or len(self._slot_steps) != len(self._slots):
            self._slot_steps = [self.t] * len(self._slots)
        late = [i for i, s in enumerate(self._slots) if s[0].grad is not None and self._slot_steps[i] != self.t]
        if late:
            raise RuntimeError("MetaTrainerStep: %d parameters (e.g. %s) receive a gradient now but skipped earlier steps; "
                               "torch.optim.Adam would bias-correct them with their own step count (%d), the fused step "
                               "has one count (%d) per weight buffer" % (len(late), [self._slots[i][5] for i in late[:3]],
                                                                          self._slot_steps[late[0]] + 1, self.t + 1))
        for i, s in enumerate(self._slots):
            if s[0].grad is not None:
                self._slot_steps[i] = self.t + 1
        last = getattr(self.model, "last_meta_grads", None)
        G = last["all"] if last is not None else None
        ok = G is not None
        gbase = G.data_ptr() if ok else 0
        holes = []
        for s in self._slots:
            g = s[0].grad
            if g is None:
                holes.append(s)
            elif ok and g.data_ptr() != gbase + s[2]:
                ok = False
                break
        if not ok:
            if self._own_g is None:
                self._own_g = self.loop.ops.zeros(1, sum(self._sizes()))
            G = self._own_g
            _mark_no_grad(G)
            for p, _, _, base, pack, name in self._slots:
                if p.grad is not None:
                    pack.view(G[:, base:base + pack.numel], name)[0].copy_(p.grad)
            return G
        for _, _, _, base, pack, name in holes:      # .grad is None: clip_grad_norm_ and Adam skip the parameter
            _mark_no_grad(pack.view(G[:, base:base + pack.numel], name))
        return G

    def _rebuild_twins(self):
        """W^T twins (and TF32 twins in single-pass mode) of the updated weights.  ~110 transposes on
        static addresses: captured once into a CUDA graph and replayed."""
        lp = self.loop
        use_graph = getattr(self.model, "use_cuda_graph", False) and getattr(lp.ops, "device", None) is not None \
            and torch.device(lp.ops.device).type == "cuda"
        if not use_graph:
            lp.refresh_weights(repack=False)
            return
        if self._twin_graph is None:
            lp.refresh_weights(repack=False)                       # warm-up outside the capture
            torch.cuda.synchronize()
            self._twin_graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._twin_graph):
                lp.refresh_weights(repack=False)
        self._twin_graph.replay()

    # ------------------------------------------------------------------ the step
    def _kernels(self, G):
        """The 4 launches: global sum of squares, then clip + Adam + zero_grad per weight buffer."""
        ops = self.loop.ops
        self.t += 1
        part = ops.sumsq_partials(G)
        base = 0
        for key, pack, params, flat in self._segments():
            n = pack.numel
            lr = self.supervisor_lr if key == "phi" else self.detector_lr
            ops.clip_adam_step_(flat.view(-1), G[base:base + n], self.m[base:base + n], self.v[base:base + n], part,
                                self.max_norm, lr, self.betas, self.eps, self.t, zero_grad=True,
                                norm_out=self.norm if base == 0 else None)
            base += n

    # ------------------------------------------------------------------ resume
    def state_dict(self):
        """Optimiser state for resuming: Adam moments on the flat [theta | psi | phi] layout, the step count, the
        LR-schedule position and the per-parameter step counts (the weights themselves are the model's)."""
        return {"m": self.m.detach().clone(), "v": self.v.detach().clone(), "t": self.t, "tokens": self.tokens,
                "supervisor_lr": self.supervisor_lr, "sizes": list(self._sizes()),
                "slot_steps": None if self._slot_steps is None else list(self._slot_steps)}

    def load_state_dict(self, sd):
        if list(sd["sizes"]) != list(self._sizes()):
            raise ValueError(f"optimizer state was saved for flat buffers {sd['sizes']}, this model has {self._sizes()}")
        self.m.copy_(sd["m"].to(self.m.device))
        self.v.copy_(sd["v"].to(self.v.device))
        self.t, self.tokens, self.supervisor_lr = int(sd["t"]), sd["tokens"], float(sd["supervisor_lr"])
        self._slot_steps = None if sd.get("slot_steps") is None else list(sd["slot_steps"])

    def step(self, n_frames=0):
        """clip + both Adam steps + zero_grad.  n_frames = batch * frames of this iteration (drives the
        optional supervisor LR schedule, reference :113-124).  -> {"grad_norm": 0-d tensor, "lr": float}"""
        model = self.model
        if not self._aliased():             # model.to(...) / new tensors behind the Parameters since the last step
            self._alias_parameters()
        self._kernels(self._flat_grads().view(-1))
        for p in self._all_params:
            p.grad = None
        self._rebuild_twins()
        model._loop_key = model._weights_key()
        used_lr = self.supervisor_lr
        if self.lr_decay:
            self.tokens += n_frames
            if self.tokens < self.warmup_tokens:
                mult = float(self.tokens) / float(max(1, self.warmup_tokens))
            else:
                progress = float(self.tokens - self.warmup_tokens) / float(max(1, self.final_tokens - self.warmup_tokens))
                mult = max(0.1, 0.5 * (1.0 + math.cos(math.pi * progress)))
            self.supervisor_lr = self.base_supervisor_lr * mult
        return {"grad_norm": self.norm[0], "lr": used_lr}


class CheckpointAverager:
    """The reference trainer's windowed checkpoint average (engine/interactron_trainer.py:48-65,161-163): over the
    last SAVE_WINDOW epochs `record_checkpoint(w = 1/SAVE_WINDOW)` accumulates `w * state_dict()`, and
    `save_checkpoint()` writes `{"model": accumulated}` (or the plain state_dict if nothing was recorded).

    Here the trainable weights live in three flat buffers (theta | psi | phi; the Parameters alias them after
    `MetaTrainerStep`), so one record is ONE `itn_ckpt_accumulate` launch per buffer into three shadow buffers
    instead of ~330 multiply / add pairs; everything else in the state_dict (the frozen backbone, FrozenBatchNorm
    statistics, GPT mask buffers, `criterion.empty_weight`, ...) is constant during training and is accumulated
    with the same two torch ops the reference uses.  `state_dict()` emits the reference's key layout and order;
    values are bit-identical to the reference's arithmetic on the same snapshots (tests/test_trainer_*.py)."""

    def __init__(self, model):
        self.model = model
        self.loop = model._get_loop()
        self.n = 0
        self._acc = None          # shadow flat buffers
        self._rest = None         # state_dict entries that are not views of the flat buffers

    def _segments(self):
        lp = self.loop
        return ((lp.theta_pack, lp.theta_params, lp.theta), (lp.psi_pack, lp.psi_params, lp.psi),
                (lp.phi_pack, lp.phi_params, lp.phi))

    def _flat_slots(self):
        """{id(parameter): (segment index, pack, name)} for the parameters that live in the flat buffers."""
        model = self.model
        if model._loop_key != model._weights_key():
            self.loop.refresh_weights()                          # parameters changed outside the fused trainer step
            model._loop_key = model._weights_key()
        out = {}
        for i, (pack, params, _) in enumerate(self._segments()):
            for name, p in zip(pack.names, params):
                out[id(p)] = (i, pack, name)
        return out

    def record_checkpoint(self, w=1.0):
        ops = self.loop.ops
        slots = self._flat_slots()
        first = self._acc is None
        if first:
            self._acc = [ops.zeros(*flat.shape) for _, _, flat in self._segments()]
            self._rest = {}
        for acc, (_, _, flat) in zip(self._acc, self._segments()):
            ops.ckpt_accumulate_(acc.view(-1), flat.view(-1), w, first)
        by_id = {id(p): p for p in self.model.parameters()}
        for k, v in self.model.state_dict().items():
            if self._key_slot(k, slots, by_id) is not None:
                continue
            self._rest[k] = w * v if first else self._rest[k] + w * v
        self.n += 1

    def _key_slot(self, key, slots, by_id):
        """state_dict key -> flat slot of the Parameter behind it (None for buffers / frozen parameters)."""
        if not hasattr(self, "_param_of_key"):
            self._param_of_key = dict(self.model.named_parameters())
        p = self._param_of_key.get(key)
        return None if p is None else slots.get(id(p))

    def state_dict(self):
        """Averaged weights with the model's state_dict keys and order (plain state_dict if nothing recorded)."""
        sd = self.model.state_dict()
        if self._acc is None:
            return sd
        slots = self._flat_slots()
        by_id = {id(p): p for p in self.model.parameters()}
        out = type(sd)()
        for k, v in sd.items():
            slot = self._key_slot(k, slots, by_id)
            if slot is None:
                out[k] = self._rest[k]
            else:
                i, pack, name = slot
                out[k] = pack.view(self._acc[i], name)[0].reshape(v.shape)
        return out

    def save_checkpoint(self, path):
        torch.save({"model": {k: v.detach().cpu() for k, v in self.state_dict().items()}}, path)
