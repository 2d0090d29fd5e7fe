"""Golden vectors for the loss terms, produced by the REFERENCE's own functions (utils/loss_utils.py l1_loss / ssim,
scene/regulation.py compute_plane_smoothness composed as in scene/gaussian_model.py:538-577).  Run in the build container:

    python oracle/make_golden_loss.py        -> tests/golden/loss_ref.npz

``lpips`` (imported at the top of utils/loss_utils.py) and ``matplotlib`` (scene/regulation.py) are not installed: both are
stubbed, neither is used by the functions called here.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))


def load_reference_loss_modules():
    from oracle import ref_loader as rl
    if not rl.reference_available():
        raise RuntimeError("reference tree not present")
    rl._install_shims()
    for name in ("lpips", "matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if "matplotlib" in sys.modules and not hasattr(sys.modules["matplotlib"], "pyplot") and "matplotlib.pyplot" in sys.modules:
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if "utils" not in sys.modules or not hasattr(sys.modules["utils"], "__path__"):
        pkg = types.ModuleType("utils")
        pkg.__path__ = [os.path.join(rl.REF_ROOT, "utils")]
        sys.modules["utils"] = pkg
    import importlib
    lu = importlib.import_module("utils.loss_utils")
    reg = importlib.import_module("scene.regulation")
    return lu, reg


def reference_compute_regulation(reg, grids, w_ts, w_l1, w_tv):
    """gaussian_model.py:538-577 with the reference's compute_plane_smoothness"""
    plane = sum(reg.compute_plane_smoothness(g[i]) for g in grids for i in (0, 1, 3))
    time = sum(reg.compute_plane_smoothness(g[i]) for g in grids for i in (2, 4, 5))
    l1 = sum(torch.abs(1 - g[i]).mean() for g in grids for i in (2, 4, 5))
    return w_tv * plane + w_ts * time + w_l1 * l1


def inputs(seed=0):
    g = torch.Generator().manual_seed(seed)
    img1 = torch.rand(2, 3, 37, 45, generator=g, dtype=torch.float64)
    img2 = (img1 + 0.1 * torch.randn(2, 3, 37, 45, generator=g, dtype=torch.float64)).clamp(0, 1)
    res = [7, 6, 5, 9]
    grids = []
    for mult in (1, 2):
        r = [x * mult for x in res[:3]] + [res[3]]
        lvl = []
        for (c0, c1) in ((0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)):
            base = 1.0 if c1 == 3 else 0.3
            lvl.append(base + 0.2 * torch.randn(1, 4, r[c1], r[c0], generator=g, dtype=torch.float64))
        grids.append(lvl)
    return img1, img2, grids


def main():
    lu, reg = load_reference_loss_modules()
    img1, img2, grids = inputs()
    out = {}
    a = img1.clone().requires_grad_(True)
    l = lu.l1_loss(a, img2); l.backward()
    out["l1"], out["l1_grad"] = l.detach().numpy(), a.grad.numpy()
    a = img1.float().clone().requires_grad_(True)
    s = lu.ssim(a, img2.float()); s.backward()
    out["ssim"], out["ssim_grad"] = s.detach().numpy(), a.grad.numpy()
    leaves = [[p.clone().requires_grad_(True) for p in lvl] for lvl in grids]
    w = (0.001, 0.0001, 0.0002)
    r = reference_compute_regulation(reg, leaves, *w); r.backward()
    out["reg"] = r.detach().numpy()
    for l_, lvl in enumerate(leaves):
        for k, p in enumerate(lvl):
            out["reg_grad_%d_%d" % (l_, k)] = p.grad.numpy()
    out["reg_weights"] = np.array(w)
    path = os.path.join(os.path.dirname(HERE), "tests", "golden", "loss_ref.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items() if not k.startswith("reg_grad")})


if __name__ == "__main__":
    main()
