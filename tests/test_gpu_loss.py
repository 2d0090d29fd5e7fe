"""GPU parity of the fused loss kernels against the CPU oracle (oracle/loss_ref.py, pinned to the reference's
own functions) and the committed reference goldens."""
import importlib
import os

import numpy as np
import pytest
import torch

from oracle import loss_ref as lr
from oracle.make_golden_loss import inputs
from util_scene import g4d, make_module

pytestmark = pytest.mark.gpu
losses = importlib.import_module("4dgaussians_b200.losses")
GOLD = os.path.join(os.path.dirname(__file__), "golden", "loss_ref.npz")


def test_l1_and_ssim_match_reference_goldens():
    z = np.load(GOLD)
    img1, img2, _ = inputs()
    a = img1.float().cuda().requires_grad_(True)
    l = losses.l1_loss(a, img2.float().cuda()); l.backward()
    assert abs(float(l) - float(z["l1"])) <= 2e-7
    assert np.abs(a.grad.cpu().numpy() - z["l1_grad"]).max() <= 1e-9
    a = img1.float().cuda().requires_grad_(True)
    s = losses.ssim(a, img2.float().cuda()); s.backward()
    assert abs(float(s) - float(z["ssim"])) <= 2e-6
    assert np.abs(a.grad.cpu().numpy() - z["ssim_grad"]).max() <= 2e-8 + 1e-4 * np.abs(z["ssim_grad"]).max()


@pytest.mark.parametrize("shape", [(3, 1014, 1352), (2, 3, 67, 131), (3, 5, 7), (1, 16, 16)])
def test_l1_ssim_vs_oracle_shapes(shape):
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.rand(shape, generator=g)
    y = (x + 0.2 * torch.randn(shape, generator=g)).clamp(0, 1)
    y.view(-1)[::7] = x.view(-1)[::7]                    # exact ties: sign(0) = 0 in the L1 gradient
    up = 0.37
    for fn_gpu, fn_ref, tol in ((losses.l1_loss, lr.l1_loss, 1e-6), (losses.ssim, lr.ssim, 1e-5)):
        a = x.cuda().requires_grad_(True)
        v = fn_gpu(a, y.cuda()); (v * up).backward()
        b = x.double().requires_grad_(True)
        w = fn_ref(b, y.double()); (w * up).backward()
        assert abs(float(v) - float(w)) <= tol * max(1.0, abs(float(w))), (fn_ref.__name__, float(v), float(w))
        gerr = float((a.grad.cpu().double() - b.grad).abs().max())
        assert gerr <= 1e-4 * float(b.grad.abs().max()) + 1e-12, (fn_ref.__name__, gerr)


@pytest.mark.parametrize("net", ["small128", "dynerf", "hypernerf", "dnerf"])
def test_plane_regulation_vs_oracle(net):
    mod = make_module(net, seed=3)
    w = (0.001, 0.0001, 0.0002)      # time_smoothness, l1_time_planes, plane_tv (arguments/dynerf/default.py:11-13)
    loss = losses.compute_regulation(mod, *w)
    (loss * 1.7).backward()
    grids = [[p.detach().cpu().double().contiguous().requires_grad_(True) for p in lvl] for lvl in mod.deformation_net.grid.grids]
    ref = lr.compute_regulation(grids, *w)
    (ref * 1.7).backward()
    assert abs(float(loss) - float(ref)) <= 2e-6 * max(1e-3, abs(float(ref))) + 1e-9
    for lvl_g, lvl_r in zip(mod.deformation_net.grid.grids, grids):
        for p, r in zip(lvl_g, lvl_r):
            assert p.grad is not None and p.grad.shape == p.shape
            e = float((p.grad.cpu().double() - r.grad).abs().max())
            assert e <= 1e-5 * float(r.grad.abs().max()) + 1e-12, e
    # every other parameter is untouched
    assert all(p.grad is None for n_, p in mod.named_parameters() if "grid" not in n_)


def test_accumulate_regulation_adds_into_existing_grads():
    dp = importlib.import_module("4dgaussians_b200.dp")
    mod = make_module("small128", seed=1)
    bucket = dp.FlatGradBucket(list(mod.flat_parameters()))
    bucket.flat.fill_(0.5)
    acc = torch.zeros((), device="cuda")
    w = (0.01, 0.0001, 0.0001)
    losses.accumulate_regulation(mod, *w, loss_accum=acc)
    losses.accumulate_regulation(mod, *w, loss_accum=acc)
    grids = [[p.detach().cpu().double().contiguous().requires_grad_(True) for p in lvl] for lvl in mod.deformation_net.grid.grids]
    ref = lr.compute_regulation(grids, *w); ref.backward()
    assert abs(float(acc) - 2 * float(ref)) <= 1e-5 * abs(float(ref))
    for lvl_g, lvl_r in zip(mod.deformation_net.grid.grids, grids):
        for p, r in zip(lvl_g, lvl_r):
            assert float((p.grad.cpu().double() - (0.5 + 2 * r.grad)).abs().max()) <= 1e-6
