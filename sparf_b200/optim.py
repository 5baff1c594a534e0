"""Fused clip + Adam + LR-schedule update over flat buffers.

Mirrors what the reference's trainers do after `loss.backward()` for one optimiser group
(source/training/engine/iter_based_trainer.py:128-147 `after_backward`; nerf_trainer.py:181-204 Adam(lr, betas=(0.9,
0.999)) + ExponentialLR with gamma = (lr_end / lr)^(1 / max_iter); joint_pose_nerf_trainer.py:513-549 pose warm-up):

    flat = FlatParameters([net.nerf, net.nerf_fine])      # params AND grads become views of two flat buffers
    opt_nerf = FusedAdam(flat, lr=opt.optim.lr, gamma=(opt.optim.lr_end / opt.optim.lr) ** (1 / max_iter),
                         max_norm=opt.nerf_gradient_clipping)
    ...  loss.backward();  flat.all_reduce();  opt_nerf.step();  flat.zero_grad()

The step counters live in device memory, so `step()` is two kernel launches with no host round trip and can be
captured into the same CUDA graph as the render step.
"""
from __future__ import annotations

import ctypes
from typing import Iterable, List, Optional

import torch

from . import _lib
from .distributed import FlatGradients


class FlatParameters(FlatGradients):
    """FlatGradients + the parameters themselves re-pointed at views of one flat fp32 buffer (values preserved)."""

    def __init__(self, modules: Iterable[torch.nn.Module]):
        super().__init__(modules)
        self.flat_param = torch.empty_like(self.flat)
        o = 0
        for p in self.params:
            n = p.numel()
            self.flat_param[o:o + n].copy_(p.data.reshape(-1))
            p.data = self.flat_param[o:o + n].view_as(p)
            o += n

    def zero_grad(self):
        self.flat.zero_()


class FusedAdam:
    def __init__(self, flat: FlatParameters, lr: float, betas=(0.9, 0.999), eps: float = 1e-8, gamma: float = 1.0,
                 warmup_steps: float = 0.0, max_norm: Optional[float] = None):
        self.flat = flat
        self.lr, self.betas, self.eps, self.gamma = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(gamma)
        self.warmup_steps = float(warmup_steps or 0.0)
        self.max_norm = float(max_norm) if max_norm else 0.0
        dev = flat.flat.device
        self.exp_avg = torch.zeros_like(flat.flat)
        self.exp_avg_sq = torch.zeros_like(flat.flat)
        self.steps = torch.zeros(2, dtype=torch.int64, device=dev)      # [updates taken, iterations seen]
        self.scratch = torch.zeros(4, dtype=torch.float64, device=dev)

    def step(self):
        L = _lib.lib()
        f = self.flat
        _lib.check(L.sparf_adam_step(f.flat.numel(), f.flat_param.data_ptr(), f.flat.data_ptr(), self.exp_avg.data_ptr(),
                                     self.exp_avg_sq.data_ptr(), self.steps.data_ptr(), self.scratch.data_ptr(),
                                     self.lr, self.gamma, self.warmup_steps, self.betas[0], self.betas[1], self.eps,
                                     self.max_norm, torch.cuda.current_stream().cuda_stream), "adam_step")

    def state_dict(self):
        return dict(exp_avg=self.exp_avg, exp_avg_sq=self.exp_avg_sq, steps=self.steps)

    def load_state_dict(self, sd):
        self.exp_avg.copy_(sd["exp_avg"]); self.exp_avg_sq.copy_(sd["exp_avg_sq"]); self.steps.copy_(sd["steps"])
