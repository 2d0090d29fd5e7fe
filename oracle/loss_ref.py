"""CPU restatement of the loss terms adjacent to the render path.  TEST INFRASTRUCTURE ONLY.

Each function cites the reference lines it follows; ``tests/test_oracle_loss.py`` pins them against the reference's OWN
functions (imported from /root/reference or the unmodified copies in oracle/_ref, with ``lpips`` / ``matplotlib`` stubbed)
and against golden vectors those functions produced (``tests/golden/loss_ref.npz``, made by ``oracle/make_golden_loss.py``).
Plain torch, differentiable, any dtype (fp64 for gradient goldens).
"""
from __future__ import annotations

import math
from typing import List, Sequence

import torch
import torch.nn.functional as F


def l1_loss(network_output: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
    """utils/loss_utils.py:20-21"""
    return (network_output - gt).abs().mean()


def gaussian_window_1d(window_size: int = 11, sigma: float = 1.5, dtype=torch.float32) -> torch.Tensor:
    """utils/loss_utils.py:26-28: fp32 tensor of python-double exps, divided by its fp32 sum"""
    g = torch.tensor([math.exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(window_size)], dtype=torch.float32)
    return (g / g.sum()).to(dtype)


def ssim(img1: torch.Tensor, img2: torch.Tensor, window_size: int = 11) -> torch.Tensor:
    """utils/loss_utils.py:37-66 (size_average=True): per-channel 2-D Gaussian window = outer product of the 1-D window,
    zero padding window_size // 2, C1 = 0.01^2, C2 = 0.03^2, mean of the map."""
    squeeze = img1.dim() == 3
    if squeeze:
        img1, img2 = img1.unsqueeze(0), img2.unsqueeze(0)
    ch = img1.shape[-3]
    w1 = gaussian_window_1d(window_size).unsqueeze(1)
    w2 = w1.mm(w1.t()).float().unsqueeze(0).unsqueeze(0)
    window = w2.expand(ch, 1, window_size, window_size).contiguous().to(img1.dtype)
    pad = window_size // 2
    mu1 = F.conv2d(img1, window, padding=pad, groups=ch)
    mu2 = F.conv2d(img2, window, padding=pad, groups=ch)
    mu1_sq, mu2_sq, mu1_mu2 = mu1 * mu1, mu2 * mu2, mu1 * mu2
    s1 = F.conv2d(img1 * img1, window, padding=pad, groups=ch) - mu1_sq
    s2 = F.conv2d(img2 * img2, window, padding=pad, groups=ch) - mu2_sq
    s12 = F.conv2d(img1 * img2, window, padding=pad, groups=ch) - mu1_mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    m = ((2 * mu1_mu2 + C1) * (2 * s12 + C2)) / ((mu1_sq + mu2_sq + C1) * (s1 + s2 + C2))
    return m.mean()


def plane_smoothness(t: torch.Tensor) -> torch.Tensor:
    """scene/regulation.py:22-28: squared second difference along H of a [1,C,H,W] plane, mean"""
    h = t.shape[-2]
    first = t[..., 1:, :] - t[..., :h - 1, :]
    second = first[..., 1:, :] - first[..., :h - 2, :]
    return (second * second).mean()


def compute_regulation(grids: Sequence[Sequence[torch.Tensor]], time_smoothness_weight: float, l1_time_planes_weight: float,
                       plane_tv_weight: float) -> torch.Tensor:
    """scene/gaussian_model.py:538-577: ``grids`` = ModuleList[L] of 6 planes [1,C,H,W];
    plane_tv * sum smooth(planes 0,1,3) + time_smoothness * sum smooth(planes 2,4,5) + l1 * sum mean|1 - planes 2,4,5|"""
    plane = sum(plane_smoothness(g[i]) for g in grids for i in (0, 1, 3))
    time = sum(plane_smoothness(g[i]) for g in grids for i in (2, 4, 5))
    l1 = sum((1 - g[i]).abs().mean() for g in grids for i in (2, 4, 5))
    return plane_tv_weight * plane + time_smoothness_weight * time + l1_time_planes_weight * l1
