"""Data-parallel plumbing for the one place the path has a real exchange step: the scene is replicated,
views / timestamps are sharded across ranks, and every training step ends with ONE all-reduce over a single flat fp32
gradient bucket (per-Gaussian 59 floats x N + HexPlane planes + MLP ~ 81 MB at C3), then identical optimizer steps on
every rank.  Rendering alone needs no collective.

The reference has no communication layer at all; this is new host-side logic, kept backend-agnostic
(`nccl` on GPUs, `gloo` in the CPU tests).
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence

import torch


class FlatGradBucket:
    """Owns one flat buffer and makes every parameter's ``.grad`` a view into it, so autograd accumulates straight into
    the bucket (no per-step flatten/unflatten copies) and the step's communication is a single collective.
    Channel-last plane parameters get channel-last gradient views (the layout the scatter kernel writes)."""

    def __init__(self, params: Sequence[torch.Tensor]):
        self.params: List[torch.Tensor] = list(params)
        if not self.params:
            raise ValueError("FlatGradBucket needs at least one parameter")
        dev, dt = self.params[0].device, self.params[0].dtype
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.numel, device=dev, dtype=dt)
        off = 0
        for p in self.params:
            n = p.numel()
            seg = self.flat[off:off + n]
            if p.dim() == 4 and p.is_contiguous(memory_format=torch.channels_last) and not p.is_contiguous():
                b, c, h, w = p.shape
                view = seg.view(b, h, w, c).permute(0, 3, 1, 2)
            else:
                view = seg.view(p.shape)
            p.grad = view
            off += n

    def zero_(self):
        self.flat.zero_()

    def allreduce_mean(self, dist=None, world_size: int = 1):
        """Sum over ranks then divide: the loss is a mean over the global batch of views (train.py:197-201)."""
        if dist is not None and world_size > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.div_(world_size)
        return self.flat


def shard_views(num_views: int, rank: int, world_size: int) -> List[int]:
    """Round-robin view -> rank assignment (view i goes to rank i mod G)."""
    return list(range(rank, num_views, world_size))


def allreduce_densification_stats(dist, world_size: int, grad_norm_accum: torch.Tensor, denom: torch.Tensor,
                                  max_radii2D: torch.Tensor):
    """Side-channel statistics the densifier needs to stay identical on every rank (train.py:195-196,261-262;
    scene/gaussian_model.py:521-523): sum of ||grad means2D||, visible count, max radii."""
    if dist is None or world_size <= 1:
        return
    packed = torch.cat([grad_norm_accum.reshape(-1), denom.reshape(-1).to(grad_norm_accum.dtype)])
    dist.all_reduce(packed, op=dist.ReduceOp.SUM)
    n = grad_norm_accum.numel()
    grad_norm_accum.copy_(packed[:n].view_as(grad_norm_accum))
    denom.copy_(packed[n:].view_as(denom).to(denom.dtype))
    dist.all_reduce(max_radii2D, op=dist.ReduceOp.MAX)
