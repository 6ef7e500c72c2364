"""Flat Adam: train.py:41,99 uses ``torch.optim.Adam`` with default
hyper-parameters; stock torch updates the ~14 parameter tensors with dozens of small
launches.  Here every parameter becomes a view of ONE flat fp32 buffer, gradients already
live in one flat bucket (``dp.GradBucket``), and the update is a single elementwise kernel
whose step counter lives on the device (CUDA-graph replayable)."""
from __future__ import annotations

import torch

from . import ops
from .dp import GradBucket

__all__ = ["FlatAdam"]


class FlatAdam:
    def __init__(self, module: torch.nn.Module, bucket: GradBucket, lr: float = 1e-3,
                 betas=(0.9, 0.999), eps: float = 1e-8):
        self.bucket = bucket
        self.params = bucket.params
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.empty(n, dtype=torch.float32, device=dev)
        off = 0
        with torch.no_grad():
            for p in self.params:                      # re-home every parameter in the flat buffer
                view = self.flat[off:off + p.numel()].view_as(p)
                view.copy_(p)
                p.data = view
                off += p.numel()
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.step_count = torch.zeros(1, dtype=torch.int64, device=dev)

    def step(self) -> None:
        ops.adam_step(self.flat, self.bucket.flat, self.exp_avg, self.exp_avg_sq, self.step_count,
                      self.lr, self.betas[0], self.betas[1], self.eps)

    def zero_grad(self) -> None:
        self.bucket.zero_()
