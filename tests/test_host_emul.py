"""Host build of the product's per-Gaussian device math (csrc/g4d_math.cuh) vs the oracle.
Catches logic errors in the GPU-less container; the real parity tests are the -m gpu ones."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import raster_ref as rr
from util_scene import cam_tuple, raster_inputs, synth

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_emul", "emul_math.cpp")
LIB = os.path.join(HERE, "host_emul", "libemul.so")


def _lib():
    hdr = os.path.join(os.path.dirname(HERE), "4dgaussians_b200", "csrc", "g4d_math.cuh")
    if not os.path.isfile(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(hdr)):
        subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-o", LIB, SRC], check=True)
    return C.CDLL(LIB)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("seed,theta,radius,wh,deg", [(0, 10.0, 4.0, (400, 400), 3), (1, -100.0, 1.5, (1352, 1014), 2),
                                                       (2, 77.0, 2.2, (536, 960), 1), (3, 160.0, 3.0, (33, 17), 0)])
def test_preprocess_bit_exact_vs_oracle(seed, theta, radius, wh, deg):
    lib = _lib()
    cam = synth.make_camera(theta, wh[0], wh[1], radius=radius)
    rc, _ = cam_tuple(cam, (0, 0, 0), sh_degree=deg, scale_modifier=1.0 if seed != 2 else 0.7)
    m, s, r, o, sh = [t.numpy().astype(np.float32) for t in raster_inputs(5000, seed, scale_mean=0.03)]
    ref = rr.preprocess(rc, m, s, r, o, sh)
    n = m.shape[0]
    depth = np.zeros(n, np.float32); radii = np.zeros(n, np.int32); xy = np.zeros((n, 2), np.float32)
    co = np.zeros((n, 4), np.float32); rgb = np.zeros((n, 3), np.float32); cl = np.zeros((n, 3), np.uint8)
    rect = np.zeros((n, 4), np.int32); tiles = np.zeros(n, np.uint32)
    lib.emul_preprocess(C.byref(rc), n, _p(m), _p(s), _p(r), _p(o), _p(sh), _p(depth), _p(radii), _p(xy), _p(co), _p(rgb),
                        _p(cl), _p(rect), _p(tiles))
    assert (ref.radii > 0).sum() > 100
    for a, b in ((depth, ref.depth), (radii, ref.radii), (xy, ref.xy), (co, ref.conic_op), (rgb, ref.rgb),
                 (cl, ref.clamped), (rect, ref.rect), (tiles, ref.tiles_touched)):
        assert np.array_equal(a.view(np.uint8), b.view(np.uint8))


def test_gaussian_backward_vs_oracle():
    lib = _lib()
    cam = synth.make_camera(25.0, 200, 120, radius=2.0)
    rc, _ = cam_tuple(cam, (0, 0, 0), sh_degree=3, scale_modifier=1.3)
    m, s, r, o, sh = [t.numpy().astype(np.float32) for t in raster_inputs(3000, 4, scale_mean=0.05)]
    pr = rr.preprocess(rc, m, s, r, o, sh)
    n = m.shape[0]
    rng = np.random.default_rng(0)
    g2 = rng.standard_normal((n, 2)).astype(np.float32); gc = rng.standard_normal((n, 3)).astype(np.float32)
    grgb = rng.standard_normal((n, 3)).astype(np.float32)
    want = rr.preprocess_backward(rc, m, s, r, sh, pr, g2, gc, grgb)
    gm = np.zeros((n, 3), np.float32); gs = np.zeros((n, 3), np.float32); gr = np.zeros((n, 4), np.float32)
    gsh = np.zeros((n, 16, 3), np.float32)
    lib.emul_backward(C.byref(rc), n, _p(m), _p(s), _p(r), _p(sh), _p(pr.radii), _p(pr.clamped), _p(g2), _p(gc), _p(grgb),
                      _p(gm), _p(gs), _p(gr), _p(gsh))
    for got, ref, nm in zip((gm, gs, gr, gsh), want, ("mean", "scale", "rot", "sh")):
        # fp32 evaluation vs the oracle's fp64: relative to the per-row magnitude
        denom = np.maximum(np.abs(ref).reshape(n, -1).max(axis=1), 1e-2).reshape((n,) + (1,) * (ref.ndim - 1))
        err = (np.abs(got - ref) / denom).max()
        assert err < 5e-3, (nm, err)
        assert np.abs(ref).max() > 0
