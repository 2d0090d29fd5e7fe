"""Generate golden vectors for the deformation path by running the REFERENCE's own module.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):
    python oracle/make_golden_deform.py
Writes tests/golden/deform_<cfg>.npz.  Weights are regenerated from ``random_params(cfg, seed)``
(deterministic CPU generator) so the fixtures stay small; a parameter checksum guards against RNG
drift, and the ``tiny`` fixture additionally stores every weight explicitly.

Outputs stored: the reference module's forward 5-tuple (fp32) and autograd gradients of
``sum(out_i * probe_i)`` w.r.t. xyz, w0 and plane (level 0, plane 2) -- enough to pin forward and
backward of the restatement.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import deform_ref as dr  # noqa: E402
from oracle.ref_loader import load_reference_deform_network  # noqa: E402


def synth_inputs(n, seed, dtype=torch.float32):
    g = torch.Generator().manual_seed(1000 + seed)
    xyz = (torch.rand(n, 3, generator=g, dtype=torch.float64) * 2 - 1) * 1.4   # some points outside the aabb
    scales = torch.log(0.02 * torch.exp(0.5 * torch.randn(n, 3, generator=g, dtype=torch.float64)))
    rot = torch.randn(n, 4, generator=g, dtype=torch.float64)
    op = torch.logit(0.05 + 0.9 * torch.rand(n, 1, generator=g, dtype=torch.float64))
    shs = torch.randn(n, 16, 3, generator=g, dtype=torch.float64) * 0.2
    probes = [torch.randn(n, 3, generator=g, dtype=torch.float64), torch.randn(n, 3, generator=g, dtype=torch.float64),
              torch.randn(n, 4, generator=g, dtype=torch.float64), torch.randn(n, 1, generator=g, dtype=torch.float64),
              torch.randn(n, 16, 3, generator=g, dtype=torch.float64)]
    return [t.to(dtype) for t in (xyz, scales, rot, op, shs)], [p.to(dtype) for p in probes]


def param_checksum(prm: dr.DeformParams) -> float:
    return float(sum(p.double().abs().sum() for p in prm.leaves()))


def main():
    out_dir = os.path.join(os.path.dirname(HERE), "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, n, seed, times in (("tiny", 61, 3, (0.0, 0.37, 1.0)), ("dnerf", 97, 1, (0.0, 0.61)),
                                 ("hypernerf", 53, 4, (0.45,)), ("dynerf", 97, 2, (0.3, 1.0))):
        cfg = dr.CONFIGS[name]
        aabb = torch.tensor([[1.31, 1.27, 1.3], [-1.29, -1.3, -1.22]])
        prm = dr.random_params(cfg, seed=seed, aabb=aabb)
        net = load_reference_deform_network(cfg)
        sd = net.state_dict()
        sd.update(dr.params_to_state_dict(prm))
        net.load_state_dict(sd)
        (xyz, sc, rot, op, shs), probes = synth_inputs(n, seed)
        blob = {"n": n, "seed": seed, "times": np.array(times, dtype=np.float32),
                "param_checksum": param_checksum(prm)}
        for ti, t in enumerate(times):
            xyz_r = xyz.clone().requires_grad_(True)
            for p in net.parameters():
                p.grad = None
            tsel = torch.tensor(t).repeat(n, 1)      # gaussian_renderer/__init__.py:52
            outs = net(xyz_r, sc, rot, op, shs, tsel)
            loss = sum((o * p).sum() for o, p in zip(outs, probes))
            loss.backward()
            for nm, o in zip(("pts", "scales", "rot", "opacity", "shs"), outs):
                blob[f"t{ti}_{nm}"] = o.detach().numpy()
            blob[f"t{ti}_g_xyz"] = xyz_r.grad.numpy()
            blob[f"t{ti}_g_w0"] = net.deformation_net.feature_out[0].weight.grad.numpy().copy()
            blob[f"t{ti}_g_plane02"] = net.deformation_net.grid.grids[0][2].grad.numpy().copy()
            blob[f"t{ti}_g_plane10"] = net.deformation_net.grid.grids[1][0].grad.numpy().copy()
        if name == "tiny":
            for k, v in dr.params_to_state_dict(prm).items():
                blob["w_" + k] = v.numpy()
        path = os.path.join(out_dir, f"deform_{name}.npz")
        np.savez_compressed(path, **blob)
        print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
