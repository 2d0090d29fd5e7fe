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
        self._views: List[torch.Tensor] = []
        off = 0
        for p in self.params:
            n = p.numel()
            seg = self.flat[off:off + n]
            if p.dim() == 4 and p.is_contiguous(memory_format=torch.channels_last) and not p.is_contiguous():
                b, c, h, w = p.shape
                view = seg.view(b, h, w, c).permute(0, 3, 1, 2)
            else:
                view = seg.view(p.shape)
            self._views.append(view)
            off += n
        self.attach()

    def attach(self):
        """(Re-)install the bucket views as the parameters' ``.grad``.  Needed after ``optimizer.zero_grad(set_to_none=True)``
        (the reference does that every step, train.py:292): autograd would otherwise allocate fresh ``.grad`` tensors outside
        the bucket and the all-reduce would run over a stale buffer.  Prefer ``bucket.zero_()`` to ``zero_grad``."""
        for p, v in zip(self.params, self._views):
            p.grad = v

    def detached(self) -> List[int]:
        """indices of parameters whose ``.grad`` is no longer the bucket view (None, or a tensor autograd / the caller
        allocated elsewhere)"""
        return [i for i, (p, v) in enumerate(zip(self.params, self._views))
                if p.grad is None or p.grad.data_ptr() != v.data_ptr() or p.grad.stride() != v.stride()]

    def zero_(self):
        self.flat.zero_()

    def allreduce_mean(self, dist=None, world_size: int = 1, on_detached: str = "raise"):
        """Sum over ranks then divide: the loss is a mean over the global batch of views (train.py:197-201).

        Every parameter's ``.grad`` must still be its bucket view, otherwise the collective would reduce a stale buffer while
        the optimizer consumes un-reduced local gradients and the ranks diverge silently.  ``on_detached``: "raise"
        (default), or "adopt" = copy the stray gradients into the bucket and re-attach (e.g. after
        ``zero_grad(set_to_none=True)``).  After densify / prune replaced parameters, build a NEW bucket."""
        bad = self.detached()
        if bad:
            if on_detached != "adopt":
                raise RuntimeError(
                    "FlatGradBucket: .grad of %d parameter(s) (first: index %d) is no longer a view of the bucket -- "
                    "optimizer.zero_grad(set_to_none=True) or a parameter swap (densify / prune) detached it. Use bucket.zero_() "
                    "instead of zero_grad, call bucket.attach() before the backward, or rebuild the bucket after changing "
                    "the parameter set." % (len(bad), bad[0]))
            for i in bad:
                g = self.params[i].grad
                if g is not None:
                    self._views[i].copy_(g)
                else:
                    self._views[i].zero_()
            self.attach()
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
