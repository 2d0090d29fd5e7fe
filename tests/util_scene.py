"""Shared helpers for tests: small synthetic scenes in the rasterizer's post-activation input space."""
import math
from importlib import import_module

import numpy as np
import torch

synth = import_module("4dgaussians_b200.synth")


def raster_inputs(n, seed, scale_mean=0.08, sh_scale=0.3):
    """(means3D, scales, rots[normalised], opac, shs) as fp64 torch tensors that are exactly fp32-representable."""
    sc = synth.make_scene(n, seed=seed, scale_mean=scale_mean)
    g = torch.Generator().manual_seed(seed + 77)
    means3D = sc["xyz"].double()
    scales = torch.exp(sc["scaling"]).float().double()
    rots = torch.nn.functional.normalize(sc["rotation"], dim=-1).float().double()
    opac = torch.sigmoid(sc["opacity"]).float().double()
    shs = torch.cat([sc["features_dc"], torch.randn(n, 15, 3, generator=g) * sh_scale], dim=1).float().double()
    return means3D, scales, rots, opac, shs


def cam_tuple(camera, bg, sh_degree=3, scale_modifier=1.0):
    from oracle import raster_ref as rr
    from oracle.dense_ref import cam_dict_from
    cd = cam_dict_from(camera, bg, sh_degree, scale_modifier)
    rc = rr.make_cam(cd["H"], cd["W"], cd["tanfovx"], cd["tanfovy"], cd["view"], cd["proj"], cd["campos"], cd["bg"],
                     sh_degree, scale_modifier)
    # use the fp32-rounded tangents in the dense path too
    cd["tanfovx"] = float(np.float32(cd["tanfovx"])); cd["tanfovy"] = float(np.float32(cd["tanfovy"]))
    return rc, cd


# ---------------------------------------------------------------------------------------------------
# helpers for the GPU parity tests: oracle-side composition of the whole render() path
# ---------------------------------------------------------------------------------------------------
import importlib

g4d = importlib.import_module("4dgaussians_b200")


class OracleRaster(torch.autograd.Function):
    """CPU autograd wrapper around the C rasterizer oracle (forward + hand-derived backward)."""

    @staticmethod
    def forward(ctx, means3D, shs, opac, scales, rots, rc):
        from oracle import raster_ref as rr
        arrs = [t.detach().numpy().astype(np.float32) for t in (means3D, scales, rots, opac, shs)]
        f = rr.rasterize_forward(rc, *arrs)
        ctx.rc, ctx.f, ctx.arrs = rc, f, arrs
        return torch.from_numpy(f["color"].copy()), torch.from_numpy(f["depth"].copy()), torch.from_numpy(f["radii"].copy())

    @staticmethod
    def backward(ctx, g_color, _gd, _gr):
        from oracle import raster_ref as rr
        g = rr.rasterize_backward(ctx.rc, *ctx.arrs, ctx.f, g_color.numpy().astype(np.float32))
        ctx.means2D_grad = g["means2D"]
        OracleRaster.last_means2D_grad = g["means2D"]
        return (torch.from_numpy(g["means3D"]), torch.from_numpy(g["shs"]), torch.from_numpy(g["opacities"]),
                torch.from_numpy(g["scales"]), torch.from_numpy(g["rots"]), None)


def oracle_params_from_module(module):
    """DeformConfig + DeformParams (CPU fp32 leaves with requires_grad) mirroring a g4d deform_network."""
    from oracle import deform_ref as dr
    a = module.args
    kc = a.kplanes_config
    cfg = dr.DeformConfig(channels=kc["output_coordinate_dim"], resolution=tuple(kc["resolution"]), multires=tuple(a.multires),
                          net_width=a.net_width, no_dx=a.no_dx, no_ds=a.no_ds, no_dr=a.no_dr, no_do=a.no_do, no_dshs=a.no_dshs)
    sd = {k: v.detach().cpu().contiguous().clone() for k, v in module.state_dict().items()}
    prm = dr.params_from_state_dict(sd, cfg.levels)
    for t in prm.leaves():
        t.requires_grad_(True)
    return cfg, prm


def oracle_render(cfg, prm, scene_cpu, camera, t, bg, sh_degree=3, stage="fine", scale_modifier=1.0):
    """Reference composition (gaussian_renderer/__init__.py:80-128) on the oracle: deform -> activations -> rasterize.
    scene_cpu: dict of CPU tensors (leaves may require grad).  Returns (color, depth, radii, rc)."""
    from oracle import deform_ref as dr
    shs = torch.cat([scene_cpu["features_dc"], scene_cpu["features_rest"]], dim=1)
    if stage == "fine":
        pts, sc, rot, op, sh = dr.deform_forward(cfg, prm, scene_cpu["xyz"], scene_cpu["scaling"], scene_cpu["rotation"],
                                                 scene_cpu["opacity"], shs, float(t))
    else:
        pts, sc, rot, op, sh = scene_cpu["xyz"], scene_cpu["scaling"], scene_cpu["rotation"], scene_cpu["opacity"], shs
    s, r, o = dr.activate(sc, rot, op)
    rc, _ = cam_tuple(camera, bg, sh_degree=sh_degree, scale_modifier=scale_modifier)
    color, depth, radii = OracleRaster.apply(pts, sh, o, s, r, rc)
    return color, depth, radii, rc, (pts, s, r, o, sh)


def make_module(net: str, seed: int = 0, device="cuda", aabb=None):
    torch.manual_seed(1234 + seed)          # nn.Linear / plane initialisation draws from the global generator
    m = g4d.deform_network(synth.hidden_args(net))
    synth.perturb_deformation(m, seed)
    if aabb is not None:
        m.deformation_net.set_aabb(aabb[0].tolist(), aabb[1].tolist())
    return m.to(device)


def rel_err(got, ref, floor=1e-3):
    got = np.asarray(got, dtype=np.float64); ref = np.asarray(ref, dtype=np.float64)
    return float(np.abs(got - ref).max() / max(floor, np.abs(ref).max()))


def rel_err_bulk(got, ref, floor=1e-3, q=99.9):
    """(q-th percentile, max) of |got - ref| relative to max|ref|.  For gradients that pass through ReLU / floor()
    decisions: an input sitting within fp32 rounding of a kink may legitimately land on the other side on the GPU."""
    got = np.asarray(got, dtype=np.float64); ref = np.asarray(ref, dtype=np.float64)
    e = np.abs(got - ref).reshape(-1) / max(floor, np.abs(ref).max())
    return float(np.percentile(e, q)), float(e.max())


def kink_rows(cfg, prm, xyz, t, delta_act=2e-6, delta_grid=2e-4):
    """[N] bool: Gaussians sitting within rounding distance of a NON-DIFFERENTIABLE point of the deformation network,
    evaluated in fp64 on the oracle: a pre-activation of an active ReLU within delta_act of 0, or a HexPlane sample
    coordinate within delta_grid (in texels) of a grid line, where floor() picks the bilinear cell.  On such rows an fp32
    implementation may legitimately take the other branch, and the gradient is then the (equally valid) one-sided
    derivative of the neighbouring piece.  Everywhere else gradients must agree to the stated tolerance."""
    from oracle import deform_ref as dr
    with torch.no_grad():
        x64 = xyz.detach().double().cpu()
        n = x64.shape[0]
        planes = [[p.detach().double() for p in lvl] for lvl in prm.planes]
        aabb = prm.aabb.detach().double()
        tt = torch.full((n,), float(t), dtype=torch.float64)
        p = dr.normalize(x64, aabb)
        q = torch.cat([p, tt.reshape(-1, 1)], dim=-1)
        mask = torch.zeros(n, dtype=torch.bool)
        for l, lvl in enumerate(planes):
            for k, (c0, c1) in enumerate(dr.PLANE_AXES):
                _, _, H, W = lvl[k].shape
                for coord, size in ((q[:, c0], W), (q[:, c1], H)):
                    g = (coord + 1.0) / 2.0 * (size - 1)          # unclamped: far outside the aabb the border rule is smooth
                    mask |= ((g - torch.round(g)).abs() < delta_grid) & (g > -delta_grid) & (g < size - 1 + delta_grid)
        feat = dr.hexplane_features(planes, x64, aabb, tt)
        hidden = feat @ prm.w0.detach().double().t() + prm.b0.detach().double()
        mask |= (hidden.abs() < delta_act).any(dim=1)
        active = {"pos": not cfg.no_dx, "scales": not cfg.no_ds, "rotations": not cfg.no_dr, "opacity": not cfg.no_do,
                  "shs": not cfg.no_dshs}
        a = torch.relu(hidden)
        for name, on in active.items():
            if not on:
                continue
            w1, b1, _, _ = prm.heads[name]
            z = a @ w1.detach().double().t() + b1.detach().double()
            mask |= (z.abs() < delta_act).any(dim=1)
    return mask
