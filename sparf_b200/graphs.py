"""CUDA-graph replay of a whole training step.

The hot path is ~25 kernel launches per step behind a Python autograd tape; at ~2 ms of device work the host-side
launch/bookkeeping time is a third of the step.  The kernels take no host-side data per step (job tables travel as
kernel parameters, the coarse-to-fine progress and the ray indices live in device memory), so the whole step
(render -> loss -> backward into the flat gradient buffer) can be captured once and replayed.

    step = GraphedStep(fn, static_inputs=(ray_idx,))   # fn(*static_inputs) -> tensor or tuple of tensors
    loss = step(new_ray_idx)                           # copies into the static inputs, replays, returns static outputs

Shapes are frozen at capture time; anything that changes shape (a different ray count) needs its own GraphedStep.
"""
from __future__ import annotations

from typing import Callable, Sequence

import torch


class GraphedStep:
    def __init__(self, fn: Callable, static_inputs: Sequence[torch.Tensor] = (), warmup: int = 3):
        self.static_inputs = tuple(static_inputs)
        self.fn = fn
        # warm up on a side stream: lazy initialisation (workspaces, host caches, cuBLAS-free here) must not be captured
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn(*self.static_inputs)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        # capture on the stream the warm-up ran on: autograd's AccumulateGrad nodes of leaf parameters (e.g. a pose
        # embedding) are bound to the stream they were first used on
        with torch.cuda.graph(self.graph, stream=side):
            self.outputs = fn(*self.static_inputs)

    def __call__(self, *inputs: torch.Tensor):
        assert len(inputs) == len(self.static_inputs)
        for dst, src in zip(self.static_inputs, inputs):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.outputs
