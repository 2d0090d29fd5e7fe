"""Pins oracle/raster_ref.c (C restatement of the absent rasterizer submodule; parity UNPINNED by the
reference) against: golden vectors of the in-tree pieces of the same math (SH polynomial, Sigma, camera
matrices), closed-form cases, structural properties, and an independent dense fp64 autograd formulation."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import raster_ref as rr
from oracle.dense_ref import dense_render
from util_scene import cam_tuple, raster_inputs, synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _np(*ts):
    return [t.detach().numpy().astype(np.float32) for t in ts]


# ------------------------------------------------------------------ in-tree golden pieces
def test_synth_camera_matches_reference_camera_golden():
    z = np.load(os.path.join(GOLD, "raster_parts.npz"))
    for i in range(3):
        w, h = int(z[f"cam{i}_w"]), int(z[f"cam{i}_h"])
        cam = synth.make_camera(float(z[f"cam{i}_theta"]), w, h, fovx=float(z[f"cam{i}_fovx"]))
        assert abs(cam.FoVy - float(z[f"cam{i}_fovy"])) < 1e-12
        assert np.allclose(cam.world_view_transform.numpy(), z[f"cam{i}_wvt"], atol=1e-7)
        assert np.allclose(cam.full_proj_transform.numpy(), z[f"cam{i}_full"], atol=1e-6)
        assert np.allclose(cam.camera_center.numpy(), z[f"cam{i}_center"], atol=1e-6)


@pytest.mark.parametrize("deg", [0, 1, 2, 3])
def test_sh_colour_matches_reference_eval_sh_golden(deg):
    z = np.load(os.path.join(GOLD, "raster_parts.npz"))
    dirs, sh = z["dirs"], z["sh"]
    n = dirs.shape[0]
    cam = synth.make_camera(0.0, 64, 64)
    campos = cam.camera_center.numpy().astype(np.float64)
    # place Gaussians at campos + 3*dir so that the oracle's view direction equals `dirs`
    means = (campos[None, :] + 3.0 * dirs).astype(np.float32)
    rc, _ = cam_tuple(cam, (0, 0, 0), sh_degree=deg)
    pr = rr.preprocess(rc, means, np.full((n, 3), 0.05, np.float32), np.tile([1, 0, 0, 0], (n, 1)).astype(np.float32),
                       np.full((n, 1), 0.5, np.float32), sh.astype(np.float32))
    want = np.maximum(z[f"rgb_deg{deg}"] + 0.5, 0.0)
    vis = pr.radii > 0
    assert vis.sum() >= 5
    assert np.abs(pr.rgb[vis] - want[vis]).max() < 5e-5
    assert np.array_equal(pr.clamped[vis].astype(bool), (z[f"rgb_deg{deg}"][vis] + 0.5) < 0)


def test_cov3d_matches_reference_build_scaling_rotation_golden():
    z = np.load(os.path.join(GOLD, "raster_parts.npz"))
    q = z["quat"]; s = z["scale"]
    qn = (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)
    n = q.shape[0]
    cam = synth.make_camera(0.0, 64, 64)
    rc, _ = cam_tuple(cam, (0, 0, 0))
    means = np.zeros((n, 3), np.float32)        # origin: visible from the orbit camera
    pr = rr.preprocess(rc, means, s.astype(np.float32), qn, np.full((n, 1), 0.5, np.float32), np.zeros((n, 16, 3), np.float32))
    sig = z["sigma"]
    want = np.stack([sig[:, 0, 0], sig[:, 0, 1], sig[:, 0, 2], sig[:, 1, 1], sig[:, 1, 2], sig[:, 2, 2]], -1)
    assert np.abs(pr.cov3d - want).max() < 1e-6


# ------------------------------------------------------------------ closed-form cases
def _single(cam, bg, pos, scale, opacity, rgb_target):
    sh = np.zeros((1, 16, 3), np.float32)
    sh[0, 0] = (np.asarray(rgb_target, np.float32) - 0.5) / 0.28209479177387814
    rc, cd = cam_tuple(cam, bg, sh_degree=0)
    f = rr.rasterize_forward(rc, np.asarray([pos], np.float32), np.full((1, 3), scale, np.float32),
                             np.asarray([[1, 0, 0, 0]], np.float32), np.asarray([[opacity]], np.float32), sh)
    return rc, f


def test_single_isotropic_gaussian_alpha_profile():
    cam = synth.make_camera(0.0, 64, 64)
    bg = (0.1, 0.2, 0.3)
    rc, f = _single(cam, bg, (0, 0, 0), 0.1, 0.8, (1.0, 0.5, 0.25))
    pr = f["proj"]
    assert pr.radii[0] > 0
    cx, cy = pr.xy[0]
    conic = pr.conic_op[0]
    ys, xs = np.mgrid[0:64, 0:64].astype(np.float32)
    dx, dy = cx - xs, cy - ys
    power = -0.5 * (conic[0] * dx * dx + conic[2] * dy * dy) - conic[1] * dx * dy
    alpha = np.minimum(0.99, 0.8 * np.exp(power))
    alpha[(alpha < 1 / 255.0) | (power > 0)] = 0
    # restrict to the Gaussian's tile rect
    r = pr.rect[0]
    mask = np.zeros((64, 64), bool); mask[r[1] * 16:r[3] * 16, r[0] * 16:r[2] * 16] = True
    alpha[~mask] = 0
    for ch, c in enumerate((1.0, 0.5, 0.25)):
        want = c * alpha + (1 - alpha) * bg[ch]
        assert np.abs(f["color"][ch] - want).max() < 2e-6
    assert np.abs(f["depth"][0] - pr.depth[0] * alpha).max() < 1e-5
    # projected centre: origin is on the optical axis of the orbit camera -> image centre (pixel coords (W-1)/2)
    assert abs(cx - 31.5) < 1e-3 and abs(cy - 31.5) < 1e-3
    assert abs(pr.depth[0] - 4.0) < 1e-5


def test_two_gaussians_front_to_back_compositing():
    cam = synth.make_camera(0.0, 32, 32)
    bg = (1.0, 1.0, 1.0)
    cc = cam.camera_center.numpy()
    fwd = -cc / np.linalg.norm(cc)
    p_near, p_far = cc + fwd * 3.0, cc + fwd * 5.0
    sh = np.zeros((2, 16, 3), np.float32)
    cols = np.array([[0.9, 0.1, 0.2], [0.2, 0.8, 0.4]], np.float32)
    sh[:, 0] = (cols - 0.5) / 0.28209479177387814
    rc, _ = cam_tuple(cam, bg, sh_degree=0)
    # far Gaussian listed first: order must come from depth, not index
    f = rr.rasterize_forward(rc, np.stack([p_far, p_near]).astype(np.float32), np.full((2, 3), 0.5, np.float32),
                             np.tile([1, 0, 0, 0], (2, 1)).astype(np.float32), np.array([[0.6], [0.7]], np.float32),
                             np.stack([sh[1], sh[0]]))
    pr = f["proj"]
    y, x = 16, 16

    def alpha(i):
        dx, dy = pr.xy[i, 0] - x, pr.xy[i, 1] - y
        co = pr.conic_op[i]
        return min(0.99, co[3] * math.exp(-0.5 * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy))
    a_near, a_far = alpha(1), alpha(0)
    for ch in range(3):
        want = cols[0, ch] * a_near + cols[1, ch] * a_far * (1 - a_near) + (1 - a_near) * (1 - a_far) * bg[ch]
        assert abs(f["color"][ch, y, x] - want) < 2e-6
    assert f["n_contrib"][y, x] == 2


def test_edge_cases_culled_and_empty():
    cam = synth.make_camera(0.0, 40, 24)
    rc, _ = cam_tuple(cam, (0.3, 0.3, 0.3))
    cc = cam.camera_center.numpy()
    fwd = -cc / np.linalg.norm(cc)
    means = np.stack([cc - fwd * 1.0,          # behind the camera
                      cc + fwd * 0.15,         # closer than the 0.2 near cull
                      cc + fwd * 4 + np.array([50, 0, 0]),   # far off-screen
                      cc + fwd * 4]).astype(np.float32)
    n = 4
    f = rr.rasterize_forward(rc, means, np.full((n, 3), 0.01, np.float32), np.tile([1, 0, 0, 0], (n, 1)).astype(np.float32),
                             np.full((n, 1), 0.9, np.float32), np.zeros((n, 16, 3), np.float32))
    assert f["radii"][0] == 0 and f["radii"][1] == 0 and f["radii"][2] == 0 and f["radii"][3] > 0
    # empty scene renders the background
    e = rr.rasterize_forward(rc, np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32), np.zeros((0, 4), np.float32),
                             np.zeros((0, 1), np.float32), np.zeros((0, 16, 3), np.float32))
    assert np.allclose(e["color"], 0.3) and e["bin"].R == 0 and np.all(e["depth"] == 0)


# ------------------------------------------------------------------ structural properties of the binning
@pytest.mark.parametrize("seed", [0, 1])
def test_binning_properties(seed):
    cam = synth.make_camera(30.0 * seed, 200, 136)
    rc, _ = cam_tuple(cam, (0, 0, 0))
    m, s, r, o, sh = _np(*raster_inputs(2000, seed, scale_mean=0.03))
    pr = rr.preprocess(rc, m, s, r, o, sh)
    bn = rr.bin_instances(rc, pr)
    assert bn.R == int(pr.tiles_touched.sum())
    assert np.all(np.diff(bn.keys.astype(np.uint64)) >= 0)                  # sorted
    tile = (bn.keys >> np.uint64(32)).astype(np.int64)
    dbits = (bn.keys & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    assert np.array_equal(dbits, pr.depth[bn.ids].view(np.uint32))          # key carries the depth bits
    same = (np.diff(bn.keys.astype(np.uint64)) == 0)
    assert np.all(np.diff(bn.ids.astype(np.int64))[same] > 0)               # ties by ascending index
    gx = (200 + 15) // 16
    for t in np.unique(tile):
        b, e = bn.ranges[t]
        assert np.all(tile[b:e] == t) and (b == 0 or tile[b - 1] != t) and (e == bn.R or tile[e] != t)
    # every (gaussian, tile) pair of the rects is present exactly once
    want = sum(int((rc_[2] - rc_[0]) * (rc_[3] - rc_[1])) for rc_ in pr.rect)
    assert want == bn.R
    g = 17
    while pr.tiles_touched[g] == 0:
        g += 1
    tiles_g = sorted(tile[bn.ids == g].tolist())
    exp = sorted(y * gx + x for y in range(pr.rect[g, 1], pr.rect[g, 3]) for x in range(pr.rect[g, 0], pr.rect[g, 2]))
    assert tiles_g == exp


# ------------------------------------------------------------------ independent dense fp64 formulation
CASES = [dict(seed=3, n=140, wh=(48, 40), theta=20.0, radius=4.0, deg=3, bg=(0.2, 0.5, 0.7), scale=0.12),
         dict(seed=5, n=90, wh=(40, 56), theta=-75.0, radius=1.6, deg=2, bg=(0, 0, 0), scale=0.06),   # near plane + guard band
         dict(seed=8, n=60, wh=(33, 17), theta=140.0, radius=3.0, deg=1, bg=(1, 1, 1), scale=0.3),
         dict(seed=9, n=200, wh=(64, 32), theta=5.0, radius=4.0, deg=0, bg=(0.5, 0.5, 0.5), scale=0.2)]  # early stop


def _run_case(c, modifier=1.0):
    cam = synth.make_camera(c["theta"], c["wh"][0], c["wh"][1], radius=c["radius"])
    rc, cd = cam_tuple(cam, c["bg"], sh_degree=c["deg"], scale_modifier=modifier)
    ins = raster_inputs(c["n"], c["seed"], scale_mean=c["scale"])
    ins = [t.clone().requires_grad_(True) for t in ins]
    dense = dense_render(cd, *ins)
    f = rr.rasterize_forward(rc, *_np(*ins))
    return rc, cd, ins, dense, f


@pytest.mark.parametrize("ci", range(len(CASES)))
def test_forward_matches_dense_fp64(ci):
    rc, cd, ins, dense, f = _run_case(CASES[ci])
    assert np.array_equal(f["radii"], dense["radii"].numpy())
    assert np.array_equal(f["proj"].rect, dense["rect"].numpy())
    assert (f["radii"] > 0).sum() > 10
    assert np.abs(f["color"] - dense["color"].detach().numpy()).max() < 2e-5
    assert np.abs(f["depth"] - dense["depth"].numpy()).max() < 1e-4
    # n_contrib may differ only where a threshold (1/255, 1e-4) is hit within fp32 rounding
    assert (f["n_contrib"] != dense["n_contrib"].numpy()).mean() < 0.01


@pytest.mark.parametrize("ci", range(len(CASES)))
def test_backward_matches_dense_autograd(ci):
    c = CASES[ci]
    rc, cd, ins, dense, f = _run_case(c)
    g = torch.Generator().manual_seed(c["seed"])
    dL = torch.randn(3, c["wh"][1], c["wh"][0], generator=g, dtype=torch.float64).float().double()
    (dense["color"] * dL).sum().backward()
    grads = rr.rasterize_backward(rc, *_np(*ins), f, dL.numpy())
    names = ("means3D", "scales", "rots", "opacities", "shs")
    for t, nm in zip(ins, names):
        want = t.grad.numpy().reshape(grads[nm].shape)
        scale = max(1e-3, np.abs(want).max())
        err = np.abs(grads[nm] - want).max() / scale
        assert err < 2e-3, (nm, err)
        assert np.abs(want).max() > 0


def test_backward_scale_modifier_and_means2D_units():
    c = CASES[0]
    cam = synth.make_camera(c["theta"], c["wh"][0], c["wh"][1], radius=c["radius"])
    rc, cd = cam_tuple(cam, c["bg"], sh_degree=c["deg"], scale_modifier=1.7)
    ins = [t.clone().requires_grad_(True) for t in raster_inputs(c["n"], c["seed"], scale_mean=c["scale"])]
    off = torch.zeros(c["n"], 2, dtype=torch.float64, requires_grad=True)
    dense = dense_render(cd, *ins, pix_offset=off)
    f = rr.rasterize_forward(rc, *_np(*ins))
    dL = torch.ones(3, c["wh"][1], c["wh"][0], dtype=torch.float64)
    (dense["color"] * dL).sum().backward()
    grads = rr.rasterize_backward(rc, *_np(*ins), f, dL.numpy())
    W, H = c["wh"]
    want2d = off.grad.numpy() * np.array([0.5 * W, 0.5 * H])      # means2D.grad is in NDC units
    s = max(1e-3, np.abs(want2d).max())
    assert np.abs(grads["means2D"][:, :2] - want2d).max() / s < 2e-3
    assert np.all(grads["means2D"][:, 2] == 0)
    want = ins[1].grad.numpy()
    s = max(1e-3, np.abs(want).max())
    assert np.abs(grads["scales"] - want).max() / s < 2e-3
