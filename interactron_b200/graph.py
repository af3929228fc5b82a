"""CUDA-graph replay of the inner loop.

One episode is ~950 kernel launches with ~0.3 ms of math: launched eagerly
from Python it is >95 % launch latency.  All shapes on the path are static for a given
(episodes, frames, resolution), so the whole adapt+detect sequence — backbone (cuDNN), our
kernels, the few torch gather/zero-fill ops — is captured once into a CUDA graph and replayed.
Weights are read from persistent flat buffers (episode.InnerLoop.refresh_weights updates them in
place), so optimizer steps between calls do not invalidate a captured graph.
"""
import gc

import torch


class GraphedCall:
    """Captures fn(*static_inputs) -> dict[str, Tensor] and replays it on new inputs."""

    def __init__(self, fn, example_inputs, warmup=2):
        self.fn = fn
        # a CUDAGraph that is garbage-collected WHILE another stream capture is in progress fails the capture
        # ("operation not permitted when stream is capturing (function reset)"): collect dead graphs first
        gc.collect()
        self.static_in = [torch.empty_like(t) for t in example_inputs]
        for s, t in zip(self.static_in, example_inputs):
            s.copy_(t)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):            # lazy inits (cuDNN plans, func attributes) stay out of the graph
                fn(*self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_out = fn(*self.static_in)
        self.replays = 0

    def __call__(self, *inputs, clone=True):
        for s, t in zip(self.static_in, inputs):
            s.copy_(t, non_blocking=True)
        self.graph.replay()
        self.replays += 1
        if not clone:
            return self.static_out
        return {k: v.clone() for k, v in self.static_out.items()}


class PipelinedPredict:
    """predict() on HOST frames with the host->device copy hidden behind the trunk.

    The frames of a step are 335 MB at 62 episodes - 6 ms of PCIe in front of a 130 ms step when they are copied
    first and the step's graph replayed afterwards.  Here the step is three graphs on static buffers: the trunk on
    the first quarter of the frames, the trunk on the rest, and everything after the trunk.  The copies run on a
    second stream; the first trunk graph starts as soon as ITS frames have arrived and covers the transfer of the
    others.  Per-frame results do not depend on how the frames are grouped into GEMM rows, so the outputs are
    bit-identical to the single-graph path (tests/test_predict_gpu.py)."""

    def __init__(self, loop, frames, masks, keys, train=False, post_frames=(0,), first=0.25):
        gc.collect()
        dev = loop.ops.device
        E, S = frames.shape[:2]
        N = E * S
        self.loop, self.keys = loop, keys
        self.n1 = max(1, min(N - 1, int(round(N * first))))
        self.frames = torch.empty(frames.shape, dtype=frames.dtype, device=dev)
        self.masks = torch.empty(masks.shape, dtype=masks.dtype, device=dev)
        self.frames.copy_(frames)
        self.masks.copy_(masks)
        flat = self.frames.flatten(0, 1)
        self.copy_stream = torch.cuda.Stream()
        self.ev = [torch.cuda.Event(), torch.cuda.Event()]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):            # lazy inits stay out of the graphs; also sizes the feature buffer
            a = loop.trunk(flat[:self.n1])
            b = loop.trunk(flat[self.n1:])
            self.src = torch.empty((N,) + tuple(a.shape[1:]), dtype=a.dtype, device=dev)
            self.src[:self.n1].copy_(a)
            self.src[self.n1:].copy_(b)
            del a, b
            loop.adapt_detect(self.frames, self.masks, post_frames=post_frames, train=train, src=self.src)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.g1, self.g2, self.g3 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g1):
            self.src[:self.n1].copy_(loop.trunk(flat[:self.n1]))
        with torch.cuda.graph(self.g2):
            self.src[self.n1:].copy_(loop.trunk(flat[self.n1:]))
        with torch.cuda.graph(self.g3):
            out = loop.adapt_detect(self.frames, self.masks, post_frames=post_frames, train=train, src=self.src)
            self.static_out = {k: out[k] for k in keys}
        self.replays = 0

    def __call__(self, frames, masks, clone=True):
        main, cs = torch.cuda.current_stream(), self.copy_stream
        cs.wait_stream(main)                     # the previous replay has finished reading the static inputs
        dst, src = self.frames.flatten(0, 1), frames.flatten(0, 1)
        with torch.cuda.stream(cs):
            dst[:self.n1].copy_(src[:self.n1], non_blocking=True)
            self.ev[0].record(cs)
            dst[self.n1:].copy_(src[self.n1:], non_blocking=True)
            self.masks.copy_(masks, non_blocking=True)
            self.ev[1].record(cs)
        main.wait_event(self.ev[0])
        self.g1.replay()
        main.wait_event(self.ev[1])
        self.g2.replay()
        self.g3.replay()
        self.replays += 1
        if not clone:
            return self.static_out
        return {k: v.clone() for k, v in self.static_out.items()}
