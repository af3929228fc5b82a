"""Benchmark of the Interactron inner-loop hot path (adapt on a 5-frame episode + re-detect).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--episodes E] [--workload NAME]
    python bench.py --impl reference ...          # the reference's CPU path on the host cores

Metric (BASELINE.json): Interactron episodes/s (inner-loop adapt+detect).  Workload =
BASELINE.json configs[2], `interactron_random.yaml` predict(): seeded synthetic 5-frame 300x300
episodes, random-init weights (`synthetic.py`).  A step = one predict() over E episodes per GPU.
One process per GPU (torchrun for N > 1); episodes are sharded, the eval path has no collective
(SURVEY.md section 8e), so scaling is weak.  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Interactron episodes/s (inner-loop adapt+detect)"
WORKLOADS = {"interactron_random": ("interactron_random", "B"), "interactron": ("interactron", "A")}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), "measured"
    return 6650.0, 1590.0, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons with nvidia-smi during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], threading.Event()
        self.t_mark = self.t_end = None

    def mark(self):
        """Start of the timed region.  The thread is started a little earlier (during the last warm-up
        steps, same load) because one nvidia-smi query takes 0.3-1 s on a multi-GPU box."""
        self.t_mark = time.perf_counter()

    def finish(self):
        """End of the timed region: stop, and let a query that is in flight (it overlaps the region) complete."""
        self.t_end = time.perf_counter()
        self.stop_flag.set()
        if self.is_alive():
            self.join(timeout=6.0)

    def run(self):
        while not self.stop_flag.is_set():
            try:
                t = time.perf_counter()
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 6:
                    self.samples.append((t, time.perf_counter(), parts))
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        timed = [p for (t0, t1, p) in self.samples if self.t_mark is None or t1 >= self.t_mark]
        window = "timed region"
        if not timed and self.samples:          # region shorter than one query: the warm-up steps just before it
            timed, window = [p for (_, _, p) in self.samples], "last warm-up steps + timed region"
        if not timed:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(s[0]) for s in timed)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in timed)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(timed[0][1]), "reasons": reasons,
                "samples": len(sm), "window": window}


def make_batch(E, first_id, pin):
    from interactron_b200.synthetic import collate_episodes, synthetic_episode
    data = collate_episodes([synthetic_episode(first_id + i, with_targets=False) for i in range(E)])
    if pin:
        data["frames"] = data["frames"].pin_memory()
        data["masks"] = data["masks"].pin_memory()
    return data


# ---------------------------------------------------------------------------------- CPU arm
def make_cpu_runner(workload):
    """Reference CPU path on the host cores.  Uses the unmodified reference when it is present
    (build container), else its pinned restatement oracle/port.py (GPU box).
    -> (run(data), kind, cores)"""
    import torch
    import interactron_b200 as ib
    from oracle import port
    from oracle import reference_harness as rh
    name, kind = WORKLOADS[workload]
    torch.set_num_threads(os.cpu_count())
    model = ib.build_model(ib.default_config(name, weights="synthetic").MODEL).eval()
    if rh.reference_available():
        ref = rh.build_reference_model(name, model.state_dict())
        return (lambda d: ref.predict(d)), "reference", torch.get_num_threads()
    sd, body = model.state_dict(), model.detector.backbone[0].body
    lr = model.config.ADAPTIVE_LR
    return (lambda d: port.predict(sd, body, d, kind, lr=lr)), "port", torch.get_num_threads()


def time_cpu(run, n_episodes, first_id=0):
    from interactron_b200.synthetic import synthetic_episode
    eps = [synthetic_episode(first_id + i, with_targets=False) for i in range(n_episodes)]
    t0 = time.perf_counter()
    for d in eps:
        run(d)
    return time.perf_counter() - t0


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = max(1, args.cpu_episodes)
    run, kind, cores = make_cpu_runner(args.workload)
    step_times = []
    for s in range(args.warmup + args.steps):
        dt = time_cpu(run, n, first_id=10 * s)
        if s >= args.warmup:
            step_times.append(dt)
    value = n * len(step_times) / sum(step_times)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "episodes/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(step_times) / len(step_times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": f"{args.workload}.yaml predict() = BASELINE configs[2] (inner-loop adapt+detect), "
                               "reference algorithm on the host CPU",
                   "episodes_per_step": n, "frames": 5, "resolution": 300, "mode": "D1 (backbone frozen)"},
        "cpu_baseline": {"value": value, "unit": "episodes/s", "cores": cores, "kind": kind,
                         "sample": f"{n} synthetic episode(s) per step, {args.steps} timed steps"},
        "e2e": {"value": value, "unit": "episodes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)



# ---------------------------------------------------------------------------------- eager PyTorch on the GPU
def eager_gpu_baseline(model, workload, n_episodes, local):
    """The reference algorithm as the reference runs it - plain torch ops + autograd, one episode at a time
    (oracle/port.py: the pinned restatement that travels; the reference tree does not) - on THIS GPU, fp32 with
    TF32 off.  The honest "beat this" line of SURVEY.md section 8d: same device, same episodes, eager launches."""
    import torch
    from oracle import port
    from interactron_b200.synthetic import synthetic_episode
    _, kind = WORKLOADS[workload]
    dev = torch.device("cuda", local)
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        sd = {k: v.detach() for k, v in model.state_dict().items()}
        body = model.detector.backbone[0].body
        lr = model.config.ADAPTIVE_LR

        def one(i):
            d = synthetic_episode(i, with_targets=False)
            d = {"frames": d["frames"].to(dev, non_blocking=True), "masks": d["masks"].to(dev, non_blocking=True)}
            out = port.predict(sd, body, d, kind, lr=lr)
            return out["pred_logits"].cpu()

        sampler = ClockSampler(local)
        sampler.start()
        for i in range(2):
            one(900 + i)
        torch.cuda.synchronize()
        sampler.mark()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n_episodes):
            one(i)
        e1.record()
        torch.cuda.synchronize()
        sampler.finish()
        ms = e0.elapsed_time(e1)
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
    return {"value": n_episodes / (ms * 1e-3), "unit": "episodes/s", "ms_per_episode": ms / n_episodes,
            "what": "oracle/port.py predict() (torch eager ops + autograd, batch 1 as the reference evaluator, fp32, "
                    "TF32 off, cuDNN backbone) on cuda:%d, host frames in / logits out" % local,
            "sample": f"{n_episodes} synthetic episodes after 2 warm-up", "clocks": sampler.summary()}


# ---------------------------------------------------------------------------------- GPU arm
def gemm_roofline(loop, frames, masks, bf16_peak):
    """One eager, instrumented step: CUDA-event pair around every tensor-core GEMM launch."""
    import torch
    ops = loop.ops
    rec = []
    att = []
    orig = ops.matmul
    orig_fwd, orig_bwd, orig_i2c = ops.attention_fwd, ops.attention_bwd, ops.im2col_nhwc
    conv_src = {}          # im2col matrix (data_ptr) -> elements of the activation it was gathered from

    def i2c(x, *a, **kw):
        out = orig_i2c(x, *a, **kw)
        conv_src[out[0].data_ptr()] = x.numel()
        return out

    def att_fwd(q, k, v, nh, scale, kmask=None, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = orig_fwd(q, k, v, nh, scale, kmask, **kw)
        e1.record()
        B, Lq, D = q.shape
        att.append((4.0 * B * Lq * k.shape[1] * D, e0, e1, 4.0 * (2 * q.numel() + 2 * k.numel())))
        return out

    def att_bwd(dO, q, k, v, *a, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = orig_bwd(dO, q, k, v, *a, **kw)
        e1.record()
        B, Lq, D = q.shape
        # 5 products (S, dP, dV, dK, dQ); operands q,k,v,o,dO read, dq,dk,dv written
        att.append((10.0 * B * Lq * k.shape[1] * D, e0, e1, 4.0 * (4 * q.numel() + 4 * k.numel())))
        return out

    def timed(a, b, **kw):
        M, K = a.shape[-2], a.shape[-1]
        N = b.shape[-1]
        nb = 1
        for x, y in zip(([1, 1] + list(a.shape[:-2]))[-2:], ([1, 1] + list(b.shape[:-2]))[-2:]):
            nb *= max(x, y)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = orig(a, b, **kw)
        e1.record()
        # algorithmic bytes: every operand once (broadcast operands once per launch), the output once,
        # plus the streamed epilogue operand (residual / mask / accumulate) when there is one
        extra = sum(1 for k in ("residual", "aux") if kw.get(k) is not None) + (1 if kw.get("accumulate") else 0)
        nbytes = 4.0 * (a.numel() + b.numel() + (1 + extra) * M * N * nb)
        # fused formulation: a convolution reads its input activation, not a materialised im2col matrix
        fused = nbytes - 4.0 * (a.numel() - conv_src[a.data_ptr()]) if a.data_ptr() in conv_src else nbytes
        rec.append((2.0 * M * N * K * nb, e0, e1, nbytes, fused))
        return out

    orig_conv = ops.conv_gemm

    def conv(x, w, kh, kw, stride, pad, dil, **kw_):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = orig_conv(x, w, kh, kw, stride, pad, dil, **kw_)
        e1.record()
        y = out[0]
        M, N = y.shape
        K = kh * kw * x.shape[-1]
        extra = 1 if kw_.get("residual") is not None else 0
        nbytes = 4.0 * (x.numel() + N * K + (1 + extra) * M * N)       # implicit GEMM: the activation itself is the A operand
        rec.append((2.0 * M * N * K, e0, e1, nbytes, nbytes))
        return out

    ops.matmul = timed
    ops.conv_gemm = conv
    ops.attention_fwd, ops.attention_bwd, ops.im2col_nhwc = att_fwd, att_bwd, i2c
    try:
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ops.launch_count()
        # a device-side delay lets the host queue run ahead, so that the per-launch event pairs
        # bracket kernel execution and not the Python launch gaps of this eager step
        torch.cuda._sleep(int(0.05 * 1.9e9))
        t0.record()
        loop.adapt_detect(frames, masks, post_frames=(0,))
        t1.record()
        torch.cuda.synchronize()
        launches = ops.launch_count() - l0
    finally:
        ops.matmul = orig
        ops.conv_gemm = orig_conv
        ops.attention_fwd, ops.attention_bwd, ops.im2col_nhwc = orig_fwd, orig_bwd, orig_i2c
    flops = sum(r[0] for r in rec)
    ms = sum(r[1].elapsed_time(r[2]) for r in rec)
    achieved = flops / (ms * 1e-3) / 1e12
    peak = bf16_peak / 2.0          # kind::tf32 runs at half the bf16 rate
    roof = {"bound": "tensor", "kernel": "gemm_tf32_kernel (tcgen05 kind::tf32, tf32x3 mode = 3 MMAs per k-step)",
            "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None,
            "mma_issue_frac": 3.0 * achieved / peak,
            "gemm_launches_per_step": len(rec), "algorithmic_gflop_per_step": flops / 1e9,
            "algorithmic_gemm_gb_per_step": sum(r[3] for r in rec) / 1e9,
            "algorithmic_gemm_gb_per_step_fused": sum(r[4] for r in rec) / 1e9,
            "gemm_ms_per_step_eager": ms, "step_ms_eager": t0.elapsed_time(t1)}
    if att:
        a_ms = sum(r[1].elapsed_time(r[2]) for r in att)
        a_fl = sum(r[0] for r in att)
        roof["attention"] = {"kernel": "attn_*_kernel (fused tcgen05 attention, tf32x3; scores never leave the SM)",
                             "launches_per_step": len(att), "ms_per_step_eager": a_ms,
                             "algorithmic_gflop_per_step": a_fl / 1e9, "achieved_tflops": a_fl / (a_ms * 1e-3) / 1e12,
                             "frac_of_tf32_peak": a_fl / (a_ms * 1e-3) / 1e12 / peak,
                             "algorithmic_gb_per_step": sum(r[3] for r in att) / 1e9}
    # DRAM traffic of the same kernel family over one step, from the committed ncu launch list of
    # `tools/profile_step.py E interactron_random` (dram__bytes_read.sum + dram__bytes_write.sum), same E as this run
    import glob
    cands = sorted(glob.glob(os.path.join(ROOT, "profiles", f"r0[2-9]*_traffic_e{frames.shape[0]}_interactron_random.json")))
    if cands and loop.kind == "B":
        tp = cands[-1]
        prof = json.load(open(tp))
        k = prof["kernels"].get("itn::gemm_tf32_kernel")
        if k:
            roof["traffic"] = (k["dram_read_bytes"] + k["dram_write_bytes"]) / k["launches"]
            roof["traffic_note"] = (f"mean DRAM bytes per GEMM launch (ncu, profiles/{os.path.basename(tp)}: "
                                    f"{(k['dram_read_bytes'] + k['dram_write_bytes']) / 1e9:.1f} GB over {k['launches']} "
                                    f"gemm_tf32_kernel launches of ONE step cut at marker launches; whole step "
                                    f"{prof['dram_gb_per_step']:.1f} GB over {prof['launches']} launches); compare with "
                                    "algorithmic_gemm_gb_per_step (every operand and output once) and ..._fused "
                                    "(convolutions read their input activation, not an im2col matrix)")
    return roof, launches


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    rank, local, world = _dist_setup()
    import interactron_b200 as ib
    name, _ = WORKLOADS[args.workload]
    model = ib.build_model(ib.default_config(name, weights="synthetic").MODEL).to(f"cuda:{local}").eval()
    loop = model._get_loop()
    E = args.episodes
    hbm_peak, bf16_peak, peak_src = load_peaks()

    batches = [make_batch(E, 1000 * rank + 100 * i, pin=True) for i in range(2)]
    dev_batches = [(b["frames"].cuda(non_blocking=True), b["masks"].cuda(non_blocking=True)) for b in batches]
    torch.cuda.synchronize()

    # instrumented eager step: per-launch GEMM timing (roofline) and launch count
    loop.adapt_detect(*dev_batches[0], post_frames=(0,))
    roof, launches_per_step = gemm_roofline(loop, *dev_batches[0], bf16_peak)
    roof["peak_source"] = f"{peak_src} bf16 dense / 2"

    keys = ("pred_logits", "pred_boxes", "image_features", "embedded_memory_features", "box_features")

    def run(f, m):
        out = loop.adapt_detect(f, m, post_frames=(0,))
        return {k: out[k] for k in keys}

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput (`value`)
    sampler = ClockSampler(local)
    for i in range(args.warmup):
        if i == args.warmup - 2:
            sampler.start()
        model._graphed("bench", run, *dev_batches[i % 2], clone=False)
    barrier()
    sampler.mark()
    ev = []
    for i in range(args.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        model._graphed("bench", run, *dev_batches[i % 2], clone=False)
        e1.record()
        ev.append((e0, e1))
    barrier()
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)

    # ---- end-to-end through the public API with host (pinned) inputs (`e2e`)
    for i in range(max(1, args.warmup // 2)):
        o = model.predict(batches[i % 2])
        o["pred_logits"].cpu()
    barrier()
    ev2 = []
    # results are read back into pinned host buffers (what a caller that cares about latency allocates once)
    lg = torch.empty((E, 1, 50, o["pred_logits"].shape[-1]), dtype=torch.float32).pin_memory()
    bx = torch.empty((E, 1, 50, 4), dtype=torch.float32).pin_memory()
    for i in range(args.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        o = model.predict(batches[i % 2])
        lg.copy_(o["pred_logits"], non_blocking=True)
        bx.copy_(o["pred_boxes"], non_blocking=True)
        e1.record()
        torch.cuda.current_stream().synchronize()
        ev2.append((e0, e1))
    barrier()
    sampler.finish()
    e2e_ms = sum(a.elapsed_time(b) for a, b in ev2)
    # predict() ships the frames as they are and, of the int64 padding masks, only the 19x19 pixels per
    # frame that the nearest-neighbour down-sampling reads (episode.sample_masks_host), as uint8
    from interactron_b200.episode import sample_masks_host
    h2d = batches[0]["frames"].numel() * 4 + sample_masks_host(batches[0]["masks"]).numel()
    d2h = lg.numel() * 4 + bx.numel() * 4

    dev_ms, e2e_ms = _max_over_ranks([dev_ms, e2e_ms], world)
    clocks = sampler.summary()
    # ---- the honest competitor on the same device: the reference algorithm in eager PyTorch (rank 0, N = 1 only)
    eager = None
    if world == 1 and args.eager_episodes > 0:
        eager = eager_gpu_baseline(model, args.workload, args.eager_episodes, local)
    # ---- the other BASELINE configs, each with its own clock record (all ranks take part: the meta step all-reduces)
    extras = None
    backbone_impl = loop.backbone_impl
    if args.extras:
        del loop, dev_batches
        model._graphs.clear()
        torch.cuda.empty_cache()
        extras = {}
        extras["baselines_configs_0_1"] = measure_baseline_models(rank, local, world)
        if args.workload != "interactron":
            extras["interactron_predict"] = measure_predict_extra("interactron", 16, 5, 3, rank, local, world)
        extras["rollout_config_3"] = measure_rollout(16, 4, 2, True, rank, local, world)
        e_meta = max(1, 16 // world) if world > 1 else 2
        extras["meta_interactron_config_4"] = measure_meta("interactron", e_meta, 8, 3, 0, rank, local, world)
    if rank == 0:
        total_eps = E * world * args.steps
        cpu = None
        if world == 1 and args.cpu_episodes > 0:
            crun, kind, cores = make_cpu_runner(args.workload)
            time_cpu(crun, 1, first_id=900)
            dt = time_cpu(crun, args.cpu_episodes)
            v = args.cpu_episodes / dt
            cpu = {"value": v, "unit": "episodes/s", "cores": cores, "kind": kind,
                   "sample": f"{args.cpu_episodes} synthetic episodes after 1 warm-up, {dt:.1f} s"}
        line = {
            "metric": METRIC, "value": total_eps / (dev_ms * 1e-3), "unit": "episodes/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "tf32x3 (fp32 storage; error-compensated 3-pass TF32 tcgen05 GEMMs, backbone convolutions included)",
            "data": "synthetic",
            "config": {"workload": f"{args.workload}.yaml predict() = BASELINE configs[2] (inner-loop adapt+detect)",
                       "episodes_per_step_per_gpu": E, "frames": 5, "resolution": 300,
                       "mode": "D1 (backbone frozen, features once per episode)", "cuda_graph": True,
                       "backbone": backbone_impl + (" (tf32x3 GEMM kernels; 3x3 / strided convolutions as implicit GEMMs through "
                                                    "TMA im2col tensor maps)" if backbone_impl == "gemm" else " (cuDNN fp32)"),
                       "l2": "256 MiB buffer written between timed steps (L2 flush); per-step CUDA events"},
            "e2e": {"value": total_eps / (e2e_ms * 1e-3), "unit": "episodes/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps,
                    "api": "model.predict(data) with pinned host tensors; logits+boxes read back into pinned host buffers"},
            "gpu_launches": launches_per_step * args.steps * 2,
            "gpu_launches_per_step": launches_per_step,
            "roofline": roof, "cpu_baseline": cpu, "eager_gpu_baseline": eager, "clocks": clocks,
            "extras": extras,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _dist_setup():
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback "
                         "(use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, local, world


def _barrier(world):
    import torch
    import torch.distributed as dist
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def _max_over_ranks(values, world):
    import torch
    import torch.distributed as dist
    t = torch.tensor(values, dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def measure_rollout(E, steps, warmup, lock, rank, local, world):
    """BASELINE configs[3]: learned-policy evaluation of `interactron.yaml` - per episode 4 policy steps
    `get_next_action()` on the 1..4 frames seen so far, then `predict()` on the 5-frame episode (reference
    engine/interactive_evaluator.py:48-66), episodes sharded over the ranks with no collective.  The
    environment is synthetic: the action does not change which frame comes next.  Everything is host-driven
    (frames arrive from the environment one at a time), so the only figure is the end-to-end one.  lock: the E
    environments of a step advance in lock-step (`get_next_actions`, one batched policy pass per step); else one
    episode at a time, batch 1, as the reference evaluator does.  -> JSON-able dict (same on every rank)."""
    import torch
    import interactron_b200 as ib
    from interactron_b200.synthetic import collate_episodes, synthetic_episode
    model = ib.build_model(ib.default_config("interactron", weights="synthetic").MODEL).to(f"cuda:{local}").eval()
    ops = model._get_ops()
    eps = []
    groups = [list(range(E))] if lock else [[e] for e in range(E)]
    for grp in groups:
        d = collate_episodes([synthetic_episode(1000 * rank + e, with_targets=False) for e in grp])
        d["frames"], d["masks"] = d["frames"].pin_memory(), d["masks"].pin_memory()
        eps.append(d)

    def view(d, s):
        o = dict(d)
        o["frames"], o["masks"] = d["frames"][:, :s], d["masks"][:, :s]
        return o

    def episode(d):
        """One rollout of the environments in `d` (all E in lock-step, or one): 4 policy steps + predict."""
        if lock:
            acts = [model.get_next_actions(view(d, s)) for s in range(1, 5)]
        else:
            acts = [model.get_next_action(view(d, s)) for s in range(1, 5)]
        out = model.predict(d)
        return acts, out["pred_logits"].cpu(), out["pred_boxes"].cpu()

    graphs_on = model.use_cuda_graph
    model.use_cuda_graph = False
    l0 = ops.launch_count()
    episode(eps[0])
    launches_per_episode = ops.launch_count() - l0
    model.use_cuda_graph = graphs_on
    sampler = ClockSampler(local)
    for i in range(warmup):
        if i == max(0, warmup - 2):
            sampler.start()
        for d in eps:
            episode(d)
    _barrier(world)
    sampler.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        for d in eps:
            episode(d)
    e1.record()
    _barrier(world)
    sampler.finish()
    ms = _max_over_ranks([e0.elapsed_time(e1)], world)[0]
    v = E * world * steps / (ms * 1e-3)
    h2d = sum(5 * 3 * 300 * 300 * 4 * s // 5 for s in (1, 2, 3, 4, 5))
    out = {
        "metric": "Interactron episodes/s (learned-policy rollout: 4x get_next_action + predict)", "value": v,
        "unit": "episodes/s", "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "tf32x3", "data": "synthetic",
        "config": {"workload": "interactron.yaml learned-policy eval = BASELINE configs[3]; host-driven, synthetic "
                               "environment; " + ("E environments advanced in lock-step (get_next_actions)" if lock
                                                  else "one episode at a time, batch 1 as in the reference"),
                   "episodes_per_step_per_gpu": E, "cuda_graph": bool(graphs_on),
                   "ms_per_episode": ms / steps / E},
        "e2e": {"value": v, "unit": "episodes/s", "h2d_bytes_per_step": h2d * E, "d2h_bytes_per_step": E * 50 * 1240 * 4,
                "api": "4x model.get_next_action(data[:s]) + model.predict(data), pinned host frames, results read back"},
        "gpu_launches": launches_per_episode * len(eps) * steps,
        "gpu_launches_per_step": launches_per_episode * len(eps),
        "roofline": None, "cpu_baseline": None, "clocks": sampler.summary()}
    del model
    torch.cuda.empty_cache()
    return out


def run_rollout_arm(args):
    """`--workload rollout`: BASELINE configs[3] on its own (also part of the default line's `extras`)."""
    import torch.distributed as dist
    rank, local, world = _dist_setup()
    out = measure_rollout(args.episodes, args.steps, args.warmup, not args.sequential, rank, local, world)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def measure_meta(name, E, steps, warmup, cpu_episodes, rank, local, world, train=True):
    """BASELINE configs[4]: the meta-training step `forward(data)` (second-order MAML gradients), the batch's
    episodes sharded over the ranks and the flat meta-gradient [theta | psi | phi] all-reduced (SUM) with NCCL in
    two buckets (the fusion part overlaps the detector pass).  The collective's time is reported separately
    (CUDA events on the streams it runs on, max over ranks).  -> JSON-able dict (same on every rank)."""
    import torch
    import interactron_b200 as ib
    from interactron_b200.synthetic import collate_episodes, synthetic_episode
    model = ib.build_model(ib.default_config(name, weights="synthetic").MODEL).to(f"cuda:{local}")
    model.train(train)            # the reference trainers run train() mode: dropout p=0.1 in every pass
    ops = model._get_ops()
    batches = []
    for i in range(2):
        d = collate_episodes([synthetic_episode(1000 * rank + 100 * i + e) for e in range(E)])
        d["frames"], d["masks"] = d["frames"].pin_memory(), d["masks"].pin_memory()
        batches.append(d)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    def step(i):
        model.zero_grad(set_to_none=True)
        _, losses = model(batches[i % 2], ridx=[(i + e) % 5 for e in range(E)])
        return float(losses["loss_supervisor_ce"])          # D2H read of the step's result

    # launches of one step, counted on an eager (non-graph) pass: graph replays bypass the counter
    graphs_on = model.use_cuda_graph
    model.use_cuda_graph = False
    l0 = ops.launch_count()
    step(0)
    torch.cuda.synchronize()
    launches_per_step = ops.launch_count() - l0
    model.use_cuda_graph = graphs_on
    sampler = ClockSampler(local)
    for i in range(warmup):
        if i == max(0, warmup - 2):
            sampler.start()
        step(i)
    _barrier(world)
    sampler.mark()
    ev, ar = [], []
    for i in range(steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step(i)
        e1.record()
        ev.append((e0, e1))
        red = model.last_meta_grads.get("allreduce")
        if red is not None:
            ar.append(red.ms())
    _barrier(world)
    sampler.finish()
    launches = launches_per_step * steps
    n_ar = max(1, len(ar))
    ms, ar_side, ar_exposed = _max_over_ranks([sum(a.elapsed_time(b) for a, b in ev), sum(a for a, _ in ar) / n_ar,
                                               sum(b for _, b in ar) / n_ar], world)
    flat = model.last_meta_grads
    n_flat, n_phi = flat["all"].numel(), flat["phi"].numel()
    cpu = None
    if rank == 0 and world == 1 and cpu_episodes > 0:
        from oracle import reference_harness as rh
        if rh.reference_available():
            torch.set_num_threads(os.cpu_count())
            ref = rh.build_reference_model(name, {k: v.cpu() for k, v in model.state_dict().items()})
            d1 = collate_episodes([synthetic_episode(7)])
            t0 = time.perf_counter()
            rh.reference_forward_with_grads(ref, d1, [2])
            dt = time.perf_counter() - t0
            cpu = {"value": 1.0 / dt, "unit": "episodes/s", "cores": torch.get_num_threads(), "kind": "reference",
                   "sample": f"1 episode of the reference forward(), {dt:.1f} s"}
    v = E * world * steps / (ms * 1e-3)
    out = {
        "metric": "Interactron episodes/s (meta-training step, second-order MAML)", "value": v,
        "unit": "episodes/s", "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "tf32x3" + (" (split accumulators: ITN_PREC_TF32X3_SPLIT)" if getattr(model, "meta_split_acc", False) else ""),
        "data": "synthetic",
        "config": {"workload": f"{name}.yaml forward() = BASELINE configs[4] (meta-training step), " +
                               ("train() mode (dropout p=0.1 in every pass, as the reference trainer)" if train
                                else "eval() mode (no dropout)") + ", D1", "episodes_per_step_per_gpu": E,
                   "global_batch": E * world, "cuda_graph": bool(model.use_cuda_graph),
                   "l2": "256 MiB buffer written between timed steps"},
        "allreduce": {"collective": "ncclAllReduce(SUM, fp32) over NVLink, 2 buckets: phi on a side stream under the "
                                    "1-frame detector backward, then theta|psi" if world > 1 else "none (single process)",
                      "elems": n_flat if world > 1 else 0, "bytes": 4 * n_flat if world > 1 else 0,
                      "bucket_elems": [n_phi, n_flat - n_phi],
                      "ms_overlapped_bucket": ar_side, "ms_exposed": ar_exposed,
                      "exposed_frac_of_step": ar_exposed / (ms / steps),
                      "timing": "CUDA events on the streams the collectives run on, mean per step, max over ranks"},
        "e2e": {"value": v, "unit": "episodes/s", "h2d_bytes_per_step": batches[0]["frames"].numel() * 4 +
                E * 5 * 361, "d2h_bytes_per_step": 4,
                "api": "model(data) with pinned host frames; grads left on .grad; one loss read back"},
        "gpu_launches": launches, "gpu_launches_per_step": launches_per_step, "roofline": None,
        "cpu_baseline": cpu, "clocks": sampler.summary()}
    del model, flush
    torch.cuda.empty_cache()
    return out


def run_meta_arm(args):
    """`--workload meta_*`: BASELINE configs[4] on its own (also part of the default line's `extras`)."""
    import torch.distributed as dist
    rank, local, world = _dist_setup()
    out = measure_meta(args.workload[len("meta_"):], args.episodes, args.steps, args.warmup, args.cpu_episodes,
                       rank, local, world, train=not args.eval_mode)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def measure_predict_extra(name, E, steps, warmup, rank, local, world):
    """predict() of another model family (device-resident CUDA-graph replay, L2 flushed between steps)."""
    import torch
    import interactron_b200 as ib
    model = ib.build_model(ib.default_config(name, weights="synthetic").MODEL).to(f"cuda:{local}").eval()
    loop = model._get_loop()
    b = make_batch(E, 1000 * rank, pin=False)
    f, m = b["frames"].cuda(), b["masks"].cuda()
    keys = ("pred_logits", "pred_boxes")

    def run(f_, m_):
        out = loop.adapt_detect(f_, m_, post_frames=(0,))
        return {k: out[k] for k in keys}

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    sampler = ClockSampler(local)
    for i in range(warmup):
        if i == max(0, warmup - 2):
            sampler.start()
        model._graphed("bench", run, f, m, clone=False)
    _barrier(world)
    sampler.mark()
    ev = []
    for i in range(steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        model._graphed("bench", run, f, m, clone=False)
        e1.record()
        ev.append((e0, e1))
    _barrier(world)
    sampler.finish()
    ms = _max_over_ranks([sum(a.elapsed_time(b_) for a, b_ in ev)], world)[0]
    out = {"metric": METRIC, "value": E * world * steps / (ms * 1e-3), "unit": "episodes/s", "n_gpus": world,
           "steps": steps, "warmup": warmup, "ms_per_step": ms / steps,
           "config": {"workload": f"{name}.yaml predict() (fusion A, 2060-token GPT; the adapt+detect half of BASELINE "
                                  "configs[3]), device-resident, CUDA graph", "episodes_per_step_per_gpu": E},
           "clocks": sampler.summary()}
    del model, loop, flush
    torch.cuda.empty_cache()
    return out


def measure_baseline_models(rank, local, world):
    """BASELINE configs[0] and [1]: the forward-only baselines `detr.predict` (1 frame) and
    `detr_multiframe.predict` (5 frames + fusion A), eager launches, host frames in / logits out."""
    import torch
    import interactron_b200 as ib
    from interactron_b200.synthetic import synthetic_episode
    res = {}
    for name, cfg, frames in (("detr", "single_frame_baseline", 1), ("detr_multiframe", "multi_frame_baseline", 5)):
        model = ib.build_model(ib.default_config(cfg, weights="synthetic").MODEL).to(f"cuda:{local}").eval()
        d = synthetic_episode(3 + rank, with_targets=False)
        d = {"frames": d["frames"][:, :frames].pin_memory(), "masks": d["masks"][:, :frames].pin_memory()}
        for _ in range(3):
            model.predict(d)["pred_logits"].cpu()
        torch.cuda.synchronize()
        n = 10
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            model.predict(d)["pred_logits"].cpu()
        e1.record()
        torch.cuda.synchronize()
        ms = _max_over_ranks([e0.elapsed_time(e1)], world)[0]
        res[name] = {"value": n * world / (ms * 1e-3), "unit": "calls/s", "ms_per_call": ms / n, "frames_per_call": frames,
                     "config": f"{cfg}.yaml {name}.predict() = BASELINE configs[{0 if frames == 1 else 1}], batch 1, eager launches, "
                               "pinned host frames in, logits read back"}
        del model
        torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--episodes", type=int, default=None,
                    help="episodes per step per GPU (default 62: 62 x 1805 encoder rows = 5.91 waves of 128-row tiles on "
                         "148 SMs, against 3.05 waves - a 4th, 5 %%-full round - at 32; 2 for meta_*)")
    ap.add_argument("--workload", default="interactron_random",
                    choices=sorted(WORKLOADS) + ["meta_" + w for w in sorted(WORKLOADS)] + ["rollout"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--sequential", action="store_true",
                    help="rollout workload: one episode at a time (batch 1) instead of E environments in lock-step")
    ap.add_argument("--cpu-episodes", type=int, default=None,
                    help="episodes in the bounded CPU-baseline sample (0 disables)")
    ap.add_argument("--eager-episodes", type=int, default=20,
                    help="episodes of the eager-PyTorch-on-GPU baseline (oracle/port.py on cuda, TF32 off; N = 1 only; 0 disables)")
    ap.add_argument("--eval-mode", action="store_true",
                    help="meta_* workloads: run forward() in eval() mode (no dropout) instead of the trainers' train() mode")
    ap.add_argument("--no-extras", dest="extras", action="store_false",
                    help="skip the `extras` object (the other BASELINE configs: baselines, interactron predict, rollout, meta step)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.cpu_episodes is None:
        args.cpu_episodes = 1 if args.impl == "reference" else 40       # ~10 s of host time
    meta = args.workload.startswith("meta_")
    if args.episodes is None:
        args.episodes = 2 if meta else (8 if args.workload == "rollout" else 62)
    if args.workload == "rollout":
        if args.impl == "reference":
            raise SystemExit("--impl reference times the predict() workloads")
        run_rollout_arm(args)
    elif meta and args.impl == "b200":
        run_meta_arm(args)
    elif args.impl == "reference":
        if meta:
            raise SystemExit("--impl reference times the predict() workloads; meta_* reports its CPU baseline inline")
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
