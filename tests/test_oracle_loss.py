"""Pins oracle/loss_ref.py (CPU restatement of l1_loss / ssim / compute_regulation) to the reference:
golden vectors produced by the reference's own functions (tests/golden/loss_ref.npz) and, when the reference tree is
present (build container) or materialised in oracle/_ref, a live comparison on fresh inputs."""
import os

import numpy as np
import pytest
import torch

from oracle import loss_ref as lr
from oracle.make_golden_loss import inputs, load_reference_loss_modules, reference_compute_regulation
from oracle.ref_loader import reference_available

GOLD = os.path.join(os.path.dirname(__file__), "golden", "loss_ref.npz")


def test_loss_oracle_matches_reference_goldens():
    z = np.load(GOLD)
    img1, img2, grids = inputs()
    a = img1.clone().requires_grad_(True)
    l = lr.l1_loss(a, img2); l.backward()
    assert abs(float(l) - float(z["l1"])) <= 1e-14 and np.abs(a.grad.numpy() - z["l1_grad"]).max() <= 1e-16
    a = img1.float().clone().requires_grad_(True)
    s = lr.ssim(a, img2.float()); s.backward()
    assert abs(float(s) - float(z["ssim"])) <= 1e-6 and np.abs(a.grad.numpy() - z["ssim_grad"]).max() <= 1e-8
    leaves = [[p.clone().requires_grad_(True) for p in lvl] for lvl in grids]
    w = tuple(float(x) for x in z["reg_weights"])
    r = lr.compute_regulation(leaves, *w); r.backward()
    assert abs(float(r) - float(z["reg"])) <= 1e-15
    for l_, lvl in enumerate(leaves):
        for k, p in enumerate(lvl):
            assert np.abs(p.grad.numpy() - z["reg_grad_%d_%d" % (l_, k)]).max() <= 1e-15, (l_, k)


@pytest.mark.skipif(not reference_available(), reason="reference tree / oracle/_ref not present")
def test_loss_oracle_matches_reference_live():
    lu, reg = load_reference_loss_modules()
    img1, img2, grids = inputs(seed=5)
    assert abs(float(lr.l1_loss(img1, img2)) - float(lu.l1_loss(img1, img2))) <= 1e-15
    assert abs(float(lr.ssim(img1.float(), img2.float())) - float(lu.ssim(img1.float(), img2.float()))) <= 1e-6
    w = (0.01, 0.0001, 0.0001)
    assert abs(float(lr.compute_regulation(grids, *w)) - float(reference_compute_regulation(reg, grids, *w))) <= 1e-15
