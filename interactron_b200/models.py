"""Drop-in model classes: same names, constructor, submodule / state_dict layout and
`predict` / `forward` / `get_next_action` / `train` / `eval` / `set_logger` surface as the
reference's models selected by MODEL.TYPE (reference utils/config_utils.py:53-77):

    interactron          reference models/interactron.py:14-197        (fusion A, learned policy)
    interactron_random   reference models/interactron_random.py:11-154 (fusion B, random policy)
    detr                 reference models/detr.py:8-84                  (single-frame baseline)
    detr_multiframe      reference models/detr_multiframe.py:9-131      (5-frame baseline, fusion A)

The arithmetic runs in the sm_100a kernels behind `ops.CudaOps`; there is no CPU / eager
fallback: calling a model whose parameters are not on a CUDA device raises.

Extensions over the reference (a strict superset of its behaviour):
  * `predict(data)` accepts b >= 1 episodes (the reference requires b == 1) and adapts them as
    one batched launch sequence; outputs keep the reference shapes, leading dim b.
  * `config.WEIGHTS == "synthetic"` builds the deterministic random-init weights of
    `synthetic.synthetic_state_dict` instead of reading a checkpoint (none is shipped offline).
"""
import os

import torch
from torch import nn

from . import modules as M
from .criterion import build_criterion
from .synthetic import synthetic_state_dict


def _load_detector_weights(detector, config, full_model):
    src = getattr(config, "WEIGHTS", None)
    if src == "synthetic":
        sd = synthetic_state_dict(full_model)
        full_model.load_state_dict(sd)
        return
    detector.load_state_dict(torch.load(src, map_location=torch.device("cpu"))["model"])


class _Base(nn.Module):
    def __init__(self):
        super().__init__()
        self.logger = None
        self.mode = "train"
        self._ops = None
        self._loop = None
        self._loop_key = None
        self._graphs = {}
        self._bb_key = None
        self._policy_cache = None           # per-frame detector outputs of the last policy step (interactron)
        self.policy_cache = True
        self.policy_cache_hits = 0
        self.use_cuda_graph = os.environ.get("ITN_CUDA_GRAPH", "1") != "0"
        self.sync_meta_grads = True
        # predict() on host frames: hide the H2D copy behind the trunk (graph.PipelinedPredict).  Off by default: measured
        # bit-identical and NOT faster on the power-capped B200 (tools/e2e_probe.py: the idle copy time is repaid by
        # higher clocks afterwards); ITN_PIPELINED_INPUT=1 enables it
        self.pipelined_input = os.environ.get("ITN_PIPELINED_INPUT", "0") != "0"
        self.meta_split_acc = os.environ.get("ITN_META_SPLITACC", "1") != "0"    # forward(): split-accumulator GEMMs
        self._drop_gen = None               # host generator of the per-step dropout seeds (train() mode)

    # -- reference surface ------------------------------------------------------------
    def eval(self):
        return self.train(False)

    def set_logger(self, logger):
        assert self.logger is None, "This model already has a logger!"
        self.logger = logger

    def get_optimizer_groups(self, train_config):
        # mirrors the reference, including its reliance on a `decoder` attribute that the
        # reference classes never define (models/interactron.py:163-168): unused by any trainer
        return [{"params": list(self.decoder.parameters()), "weight_decay": 0.0},
                {"params": list(self.detector.parameters()), "weight_decay": 0.0}]

    # -- plumbing -----------------------------------------------------------------------
    def _device(self):
        return next(self._detector().parameters()).device

    def _get_ops(self):
        if self._ops is None:
            dev = self._device()
            if dev.type != "cuda":
                raise RuntimeError("interactron_b200 models run only on a CUDA (sm_100a) device: "
                                   "call model.to('cuda') first; there is no CPU path")
            from .ops import CudaOps
            self._ops = CudaOps(dev)
        return self._ops

    def _weights_key(self):
        det, fus = self._detector(), self._fusion()
        ps = list(det.parameters()) + (list(fus.parameters()) if fus is not None else [])
        return tuple((p.data_ptr(), p._version) for p in ps)

    def _get_loop(self):
        """InnerLoop with flat weight buffers; re-packed whenever a parameter changed."""
        from .episode import InnerLoop
        ops = self._get_ops()
        key = self._weights_key()
        if self._loop is None:
            lr = getattr(self.config, "ADAPTIVE_LR", 1e-3) if hasattr(self, "config") else 1e-3
            self._loop = InnerLoop(ops, self._detector(), self._fusion(), self._kind, lr)
        elif key != self._loop_key:
            self._loop.refresh_weights()
        self._loop_key = key
        return self._loop

    def _new_dropout_seed(self, loop):
        """train() mode: draw this step's seed on the host and write it into the loop's device seed buffer (the
        kernels read it through a pointer, so replayed CUDA graphs see the new value).  The stream of seeds
        follows torch's global seed (`torch.manual_seed`, as reference train.py:15-18 sets it)."""
        if self._drop_gen is None:
            self._drop_gen = torch.Generator(device="cpu")
            self._drop_gen.manual_seed(torch.initial_seed() & 0x7FFFFFFFFFFFFFFF)
        seed = torch.randint(0, 2 ** 62, (1,), generator=self._drop_gen, dtype=torch.int64)
        if loop.drop_seed is None:
            loop.drop_seed = torch.zeros(1, dtype=torch.int64, device=loop.ops.device)
        loop.drop_seed.copy_(seed, non_blocking=True)
        cfg = getattr(self, "config", None)
        if cfg is not None and hasattr(cfg, "RESIDUAL_PDROP"):
            loop.fusion_drop = (float(cfg.RESIDUAL_PDROP), float(cfg.ATTENTION_PDROP), float(cfg.EMBEDDING_PDROP))
        return int(seed)

    def _detector(self):
        return self.detector

    def _fusion(self):
        return self.fusion

    @staticmethod
    def _frames_masks(data, device):
        """Host tensors cross PCIe here: the frames as they are; of a CPU mask only the h*w pixels per
        frame the nearest-neighbour down-sampling reads (episode.sample_masks_host)."""
        from .episode import sample_masks_host
        frames = data["frames"].to(device, non_blocking=True)
        masks = data["masks"]
        if not masks.is_cuda and torch.device(device).type == "cuda":
            masks = sample_masks_host(masks)
        masks = masks.to(device, non_blocking=True)
        return frames, masks


class _Adaptive(_Base):
    """Shared implementation of interactron / interactron_random."""

    def __init__(self, config, fusion_cls, kind):
        super().__init__()
        self.detector = M.DetectorHolder(config.NUM_CLASSES)
        self.criterion = build_criterion(config)
        self.postprocessor = {}
        self.fusion = fusion_cls(config)
        self.config = config
        self._kind = kind
        _load_detector_weights(self.detector, config, self)

    def train(self, mode=True):
        self.mode = "train" if mode else "test"
        self.detector.train(mode)
        self.fusion.train(mode)
        return self

    def predict(self, data):
        """Adapt on the episode(s) and detect on frame 0 with the adapted weights.
        Returns {k: [b,1,...]} for pred_logits, pred_boxes, image_features,
        embedded_memory_features, box_features — the reference's dict for b == 1."""
        loop = self._get_loop()
        train = self.mode == "train"         # the reference applies dropout here too when left in train() mode
        if train:
            self._new_dropout_seed(loop)
        keys = ("pred_logits", "pred_boxes", "image_features", "embedded_memory_features", "box_features")
        hf = data["frames"]
        if (self.use_cuda_graph and self.pipelined_input and not hf.is_cuda and hf.dim() == 5
                and hf.shape[0] * hf.shape[1] >= 40 and torch.device(loop.ops.device).type == "cuda"):
            # host frames, many of them: the copy runs behind the trunk (graph.PipelinedPredict)
            return self._predict_pipelined(loop, data, keys, train)
        frames, masks = self._frames_masks(data, loop.ops.device)

        def run(f, m):
            out = loop.adapt_detect(f, m, post_frames=(0,), train=train)
            return {k: out[k] for k in keys}

        if not (self.use_cuda_graph and frames.is_cuda):
            return run(frames, masks)
        return self._graphed("predict_train" if train else "predict", run, frames, masks)

    def _predict_pipelined(self, loop, data, keys, train):
        from .episode import sample_masks_host
        from .graph import PipelinedPredict
        frames, masks = data["frames"], data["masks"]
        if not masks.is_cuda:
            masks = sample_masks_host(masks)
        self._check_backbone_moved()
        key = ("predict_pipe_train" if train else "predict_pipe", tuple(frames.shape), tuple(masks.shape), masks.dtype,
               loop.ops.precision_key, loop.backbone_tf32, loop.backbone_impl, float(loop.lr), float(loop.clip),
               bool(loop.ops.fused_attention))
        g = self._graphs.get(key)
        if g is None:
            g = self._graphs[key] = PipelinedPredict(loop, frames, masks, keys, train=train)
        return g(frames, masks)

    def _check_backbone_moved(self):
        bb = self._detector().backbone
        # (data_ptr, _version): an in-place change of the frozen trunk (load_state_dict of another checkpoint)
        # re-folds BN into NEW weight tensors (backbone.refresh) the captured launches do not point to
        bb_key = tuple((t.data_ptr(), t._version) for t in list(bb.parameters()) + list(bb.buffers()))
        if bb_key != self._bb_key:          # backbone tensors moved or changed: captured conv launches are stale
            self._graphs.clear()
            self._bb_key = bb_key

    def _graphed(self, tag, fn, frames, masks, clone=True):
        """Replay (capturing on first use) the CUDA graph of `fn` for this input geometry."""
        from .graph import GraphedCall
        self._check_backbone_moved()
        # everything a capture bakes in as a kernel scalar or a branch is part of the key
        key = (tag, tuple(frames.shape), tuple(masks.shape), self._loop.ops.precision_key, self._loop.backbone_tf32,
               self._loop.backbone_impl, float(self._loop.lr), float(self._loop.clip),
               bool(self._loop.ops.fused_attention))
        g = self._graphs.get(key)
        if g is None:
            g = self._graphs[key] = GraphedCall(fn, [frames, masks])
        return g(frames, masks, clone=clone)

    def forward(self, data, train=True, ridx=None):
        """Meta-training step (reference models/interactron.py:61-151, interactron_random.py:57-136):
        -> (predictions, losses) and, as in the reference, the side effect of *accumulating* the
        batch's summed meta-gradients on `.grad` of the detector and fusion Parameters (second-order
        supervisor gradients on fusion + in_proj_*, first-order detector gradients on the fast
        weights).  `train` is ignored, as in the reference.  `ridx` (extension): the per-episode frame
        of the detector loss; default = the reference's `random.randint(0, 4)` draw per episode.
        In train() mode (what both reference trainers run, engine/interactron_trainer.py:73) dropout is applied
        in every pass with counter-based masks (layers.DropCtx); in eval() mode there is none; see meta.py."""
        from . import meta
        # one process per GPU, the batch's episodes sharded over ranks: the meta-gradient is a sum over
        # episodes -> all-reduce(SUM) of the flat buffer [theta | psi | phi] (replaces nn.DataParallel,
        # reference engine/interactron_trainer.py:43-46), in two buckets so that the fusion part overlaps
        # the detector pass (meta.meta_step); no-op in a single process
        # the meta-gradients are a second derivative through 12 post-norm layers: the step runs its GEMMs with split
        # accumulators (ITN_PREC_TF32X3_SPLIT: ~fp32 error per product, 8-20 % slower GEMMs).
        # `meta_split_acc = False` / ITN_META_SPLITACC=0: the default tf32x3 mode of predict()
        ops = self._get_loop().ops
        prev = getattr(ops, "split_acc", None)          # None: a backend without the mode (the CPU simulation in tests)
        if prev is not None:
            ops.split_acc = prev or bool(self.meta_split_acc)
        try:
            predictions, losses, flat = meta.meta_step(self, data, ridx, sync=self.sync_meta_grads)
        finally:
            if prev is not None:
                ops.split_acc = prev
        self.last_meta_grads = flat
        meta.accumulate_grads(self, flat)
        return predictions, losses


class interactron_random(_Adaptive):
    def __init__(self, config):
        super().__init__(config, M.FusionBHolder, "B")


class interactron(_Adaptive):
    def __init__(self, config):
        super().__init__(config, M.FusionAHolder, "A")
        self.path_storage = {}

    def get_next_action(self, data):
        """Policy step (reference models/interactron.py:174-197): detector + fusion A forward on the
        s frames seen so far, argmax of the s-th action head -> python int."""
        b, s = data["frames"].shape[:2]
        actions = self._policy_logits(data, 1, b * s)
        return int(actions[0, s - 1].argmax(dim=-1).item())

    def get_next_actions(self, data):
        """Extension for driving b environments in lock-step: the policy step of b independent episodes
        (each with the same number s of frames seen) in one pass -> list of b ints, element i equal to
        `get_next_action` on episode i alone.  (The reference's method folds a batch into ONE sequence of
        b*s frames, which only means something for b == 1.)"""
        b, s = data["frames"].shape[:2]
        actions = self._policy_logits(data, b, s)
        return [int(a) for a in actions[:, s - 1].argmax(dim=-1).tolist()]

    def _policy_logits(self, data, E, S):
        """Action logits [E, 4, 4] of E sequences of S frames each (E * S = all frames in `data`).

        The detector is not adapted during the rollout and its transformer treats every frame on its
        own, so the per-frame outputs the fusion network reads (encoder memory, prediction tokens) are
        kept from the previous policy step: when the first S-1 frames (and masks) are bit-identical to
        the last call's and the weights have not changed, only the new frame goes through the trunk and
        the DETR transformer (the reference recomputes all S frames each step, models/interactron.py:189).
        `policy_cache = False` switches the reuse off."""
        from . import fusion
        loop = self._get_loop()
        ops = loop.ops
        frames, masks = self._frames_masks(data, ops.device)
        frames = frames.reshape(E, S, *frames.shape[2:])
        masks = masks.reshape(E, S, *masks.shape[2:])
        graph = self.use_cuda_graph and frames.is_cuda

        def detect(f, m):
            n = f.shape[1]
            out, _, hw = loop.detect(f.flatten(0, 1), m.flatten(0, 1), want_preds=True)
            L = hw[0] * hw[1]
            return {"memory": out["memory_r"].view(E, n, L, -1), "preds": out["preds"].view(E, n, detr_nq(), -1)}

        def fuse(memory, preds):
            L = memory.shape[2]
            fout, _ = fusion.fusion_a_forward(ops, loop._fusion_weights(), memory.view(E, S * L, -1),
                                              preds.view(E * S * preds.shape[2], -1), E, S, L, need_cache=False)
            return {"actions": fout["actions"]}

        def call(tag, fn, a, b, clone):
            if not graph:
                return fn(a, b)
            return self._graphed2(tag, fn, a, b, clone)

        c = self._policy_cache
        hit = (self.policy_cache and c is not None and S > 1 and c["E"] == E and c["S"] == S - 1
               and c["wkey"] == self._loop_key and c["frames"].shape[2:] == frames.shape[2:]
               and c["masks"].shape[2:] == masks.shape[2:] and c["masks"].dtype == masks.dtype
               and torch.equal(frames[:, :S - 1], c["frames"]) and torch.equal(masks[:, :S - 1], c["masks"]))
        if hit:
            new = call(("policy_detect", E, 1), detect, frames[:, S - 1:].contiguous(), masks[:, S - 1:].contiguous(), True)
            memory = torch.cat([c["memory"], new["memory"]], 1)
            preds = torch.cat([c["preds"], new["preds"]], 1)
            self.policy_cache_hits += 1
        else:
            new = call(("policy_detect", E, S), detect, frames, masks, True)
            memory, preds = new["memory"], new["preds"]
        if self.policy_cache:
            self._policy_cache = dict(E=E, S=S, wkey=self._loop_key, frames=frames.clone(), masks=masks.clone(),
                                      memory=memory, preds=preds)
        return call(("policy_fuse", E, S), fuse, memory, preds, False)["actions"]

    def _graphed2(self, tag, fn, a, b, clone):
        """CUDA-graph replay of fn(a, b) keyed by tag and input geometry (policy rollout pieces)."""
        from .graph import GraphedCall
        self._check_backbone_moved()
        key = (tag, tuple(a.shape), tuple(b.shape), a.dtype, b.dtype, self._loop.ops.precision_key,
               bool(self._loop.ops.fused_attention))
        g = self._graphs.get(key)
        if g is None:
            g = self._graphs[key] = GraphedCall(fn, [a, b])
        return g(a, b, clone=clone)


def detr_nq():
    from . import detr_t
    return detr_t.NQ


class detr(_Base):
    """Single-frame DETR baseline (BASELINE config 1): state_dict keys `model.*`."""

    def __init__(self, config):
        super().__init__()
        self.model = M.DetectorHolder(config.NUM_CLASSES)
        self.criterion = build_criterion(config)
        self.postprocessor = {}
        self.config = config
        self._kind = "B"
        if getattr(config, "WEIGHTS", None) == "synthetic":
            self.load_state_dict(synthetic_state_dict(self))
        else:
            self.model.load_state_dict(torch.load(config.WEIGHTS, map_location=torch.device("cpu"))["model"])

    def _detector(self):
        return self.model

    def _fusion(self):
        return None

    def train(self, mode=True):
        self.mode = "train" if mode else "test"
        self.model.train(mode)
        return self

    def predict(self, data):
        """DETR forward on all b*s frames -> {k: [b,s,...]} (reference models/detr.py:20-40)."""
        loop = self._get_loop()
        frames, masks = self._frames_masks(data, loop.ops.device)
        b, s = frames.shape[:2]
        out, src_r, (h, w) = loop.detect(frames.flatten(0, 1), masks.flatten(0, 1))
        C = self.model.class_embed.out_features
        return {
            "pred_logits": out["logits"].view(b, s, 50, C),
            "pred_boxes": out["boxes"].view(b, s, 50, 4),
            "image_features": loop.src.view(b, s, h, w, -1).permute(0, 1, 4, 2, 3),
            "embedded_memory_features": out["memory"].view(b, s, h, w, -1).permute(0, 1, 4, 2, 3),
            "box_features": out["hs"].view(b, s, 50, -1),
        }

    def forward(self, data):
        raise NotImplementedError("detector training (criterion backward) is outside the inner-loop hot path")


class detr_multiframe(_Base):
    """5-frame DETR + fusion A direct-supervision baseline (BASELINE config 2)."""

    def __init__(self, config):
        super().__init__()
        self.detector = M.DetectorHolder(config.NUM_CLASSES)
        self.criterion = build_criterion(config)
        self.postprocessor = {}
        self.fusion = M.FusionAHolder(config)
        self.config = config
        self._kind = "A"
        _load_detector_weights(self.detector, config, self)

    def train(self, mode=True):
        self.mode = "train" if mode else "test"
        self.detector.train(mode)
        self.fusion.train(mode)
        return self

    def predict(self, data):
        """DETR on the b*s frames, then fusion A's box/logit decoders on one b*s-frame sequence
        (reference models/detr_multiframe.py:24-53; meaningful for b == 1 like the reference)."""
        from . import fusion
        loop = self._get_loop()
        ops = loop.ops
        frames, masks = self._frames_masks(data, ops.device)
        b, s = frames.shape[:2]
        N = b * s
        out, _, hw = loop.detect(frames.flatten(0, 1), masks.flatten(0, 1), want_preds=True)
        L = hw[0] * hw[1]
        C = self.detector.class_embed.out_features
        fout, _ = fusion.fusion_a_forward(ops, loop._fusion_weights(), out["memory_r"].view(1, N * L, -1),
                                          out["preds"], 1, N, L, need_cache=False, want_aux_heads=True)
        return {"pred_boxes": fout["pred_boxes"].view(b, s, 50, 4),
                "pred_logits": fout["pred_logits"].view(b, s, 50, C)}

    def forward(self, data):
        raise NotImplementedError("direct-supervision training is outside the inner-loop hot path")


MODEL_TYPES = {"interactron": interactron, "interactron_random": interactron_random, "detr": detr,
               "detr_multiframe": detr_multiframe}


def build_model(args):
    """Same dispatch on MODEL.TYPE as reference utils/config_utils.py:53-77 (live types only)."""
    if args.TYPE not in MODEL_TYPES:
        raise AssertionError(f"{args.TYPE} is not a valid model. Please select one from {list(MODEL_TYPES)}")
    return MODEL_TYPES[args.TYPE](args)
