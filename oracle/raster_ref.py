"""ctypes/numpy front-end of the C rasterizer oracle (oracle/raster_ref.c).

TEST INFRASTRUCTURE ONLY (see the header of raster_ref.c; parity UNPINNED by the reference because the
rasterizer submodule source and tests are absent).  Builds ``oracle/_build/libg4dref.so`` on demand
with gcc (``make -C oracle``).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libg4dref.so")
_lib = None


class RefCam(C.Structure):
    _fields_ = [("H", C.c_int32), ("W", C.c_int32), ("sh_degree", C.c_int32), ("pad_", C.c_int32),
                ("tanfovx", C.c_float), ("tanfovy", C.c_float), ("scale_modifier", C.c_float), ("pad2_", C.c_float),
                ("view", C.c_float * 16), ("proj", C.c_float * 16), ("campos", C.c_float * 3), ("bg", C.c_float * 3)]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "raster_ref.c")
    if force or not os.path.isfile(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.g4dref_count_instances.restype = C.c_int64
        assert _lib.g4dref_cam_sizeof() == C.sizeof(RefCam)
    return _lib


def make_cam(H, W, tanfovx, tanfovy, view, proj, campos, bg, sh_degree=3, scale_modifier=1.0) -> RefCam:
    cam = RefCam()
    cam.H, cam.W, cam.sh_degree = int(H), int(W), int(sh_degree)
    cam.tanfovx, cam.tanfovy, cam.scale_modifier = float(tanfovx), float(tanfovy), float(scale_modifier)
    cam.view[:] = [float(x) for x in np.asarray(view, dtype=np.float32).reshape(16)]
    cam.proj[:] = [float(x) for x in np.asarray(proj, dtype=np.float32).reshape(16)]
    cam.campos[:] = [float(x) for x in np.asarray(campos, dtype=np.float32).reshape(3)]
    cam.bg[:] = [float(x) for x in np.asarray(bg, dtype=np.float32).reshape(3)]
    return cam


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


@dataclass
class Projected:
    depth: np.ndarray; radii: np.ndarray; xy: np.ndarray; cov3d: np.ndarray; conic_op: np.ndarray
    rgb: np.ndarray; clamped: np.ndarray; rect: np.ndarray; tiles_touched: np.ndarray


def preprocess(cam: RefCam, means3D, scales, rots, opac, shs) -> Projected:
    means3D, scales, rots, opac, shs = map(_f32, (means3D, scales, rots, opac, shs))
    n = means3D.shape[0]
    out = Projected(np.zeros(n, np.float32), np.zeros(n, np.int32), np.zeros((n, 2), np.float32),
                    np.zeros((n, 6), np.float32), np.zeros((n, 4), np.float32), np.zeros((n, 3), np.float32),
                    np.zeros((n, 3), np.uint8), np.zeros((n, 4), np.int32), np.zeros(n, np.uint32))
    lib().g4dref_preprocess(C.byref(cam), n, _p(means3D), _p(scales), _p(rots), _p(opac), _p(shs), _p(out.depth),
                            _p(out.radii), _p(out.xy), _p(out.cov3d), _p(out.conic_op), _p(out.rgb), _p(out.clamped),
                            _p(out.rect), _p(out.tiles_touched))
    return out


@dataclass
class Binned:
    keys: np.ndarray; ids: np.ndarray; ranges: np.ndarray; R: int


def bin_instances(cam: RefCam, pr: Projected) -> Binned:
    n = pr.depth.shape[0]
    R = int(lib().g4dref_count_instances(n, _p(pr.tiles_touched)))
    gx, gy = (cam.W + 15) // 16, (cam.H + 15) // 16
    keys = np.zeros(max(R, 1), np.uint64); ids = np.zeros(max(R, 1), np.uint32)
    ranges = np.zeros((gx * gy, 2), np.uint32)
    lib().g4dref_bin(C.byref(cam), n, _p(pr.depth), _p(pr.rect), _p(pr.tiles_touched), C.c_int64(R), _p(keys), _p(ids),
                     _p(ranges))
    return Binned(keys[:R], ids[:R], ranges, R)


def blend_forward(cam: RefCam, pr: Projected, bn: Binned):
    H, W = cam.H, cam.W
    color = np.zeros((3, H, W), np.float32); depth = np.zeros((1, H, W), np.float32)
    final_T = np.zeros((H, W), np.float32); n_contrib = np.zeros((H, W), np.uint32)
    ids = bn.ids if bn.R > 0 else np.zeros(1, np.uint32)
    lib().g4dref_blend_forward(C.byref(cam), _p(ids), _p(bn.ranges), _p(pr.xy), _p(pr.conic_op), _p(pr.rgb),
                               _p(pr.depth), _p(color), _p(depth), _p(final_T), _p(n_contrib))
    return color, depth, final_T, n_contrib


def blend_backward(cam: RefCam, pr: Projected, bn: Binned, final_T, n_contrib, dL_dpix):
    n = pr.depth.shape[0]
    dL_dpix = _f32(dL_dpix)
    g_mean2D = np.zeros((n, 2), np.float32); g_conic = np.zeros((n, 3), np.float32)
    g_opac = np.zeros((n,), np.float32); g_rgb = np.zeros((n, 3), np.float32)
    ids = bn.ids if bn.R > 0 else np.zeros(1, np.uint32)
    lib().g4dref_blend_backward(C.byref(cam), n, _p(ids), _p(bn.ranges), _p(pr.xy), _p(pr.conic_op), _p(pr.rgb),
                                _p(final_T), _p(n_contrib), _p(dL_dpix), _p(g_mean2D), _p(g_conic), _p(g_opac),
                                _p(g_rgb))
    return g_mean2D, g_conic, g_opac, g_rgb


def preprocess_backward(cam: RefCam, means3D, scales, rots, shs, pr: Projected, g_mean2D, g_conic, g_rgb):
    means3D, scales, rots, shs = map(_f32, (means3D, scales, rots, shs))
    n = means3D.shape[0]
    g_means3D = np.zeros((n, 3), np.float32); g_scales = np.zeros((n, 3), np.float32)
    g_rots = np.zeros((n, 4), np.float32); g_shs = np.zeros((n, 16, 3), np.float32)
    lib().g4dref_preprocess_backward(C.byref(cam), n, _p(means3D), _p(scales), _p(rots), _p(shs), _p(pr.radii),
                                     _p(pr.clamped), _p(_f32(g_mean2D)), _p(_f32(g_conic)), _p(_f32(g_rgb)),
                                     _p(g_means3D), _p(g_scales), _p(g_rots), _p(g_shs))
    return g_means3D, g_scales, g_rots, g_shs


def rasterize_forward(cam: RefCam, means3D, scales, rots, opac, shs) -> Dict[str, object]:
    """Whole rasterizer forward (what GaussianRasterizer.forward returns + internals)."""
    pr = preprocess(cam, means3D, scales, rots, opac, shs)
    bn = bin_instances(cam, pr)
    color, depth, final_T, n_contrib = blend_forward(cam, pr, bn)
    return {"color": color, "depth": depth, "radii": pr.radii.copy(), "proj": pr, "bin": bn,
            "final_T": final_T, "n_contrib": n_contrib}


def rasterize_backward(cam: RefCam, means3D, scales, rots, opac, shs, fwd: Dict[str, object], dL_dcolor):
    """Gradients w.r.t. (means3D, means2D[NDC-scaled, N x 3 with z=0], shs, opacities, scales, rots)."""
    pr, bn = fwd["proj"], fwd["bin"]
    g_mean2D, g_conic, g_opac, g_rgb = blend_backward(cam, pr, bn, fwd["final_T"], fwd["n_contrib"], dL_dcolor)
    g_means3D, g_scales, g_rots, g_shs = preprocess_backward(cam, means3D, scales, rots, shs, pr, g_mean2D, g_conic, g_rgb)
    n = g_mean2D.shape[0]
    m2 = np.zeros((n, 3), np.float32); m2[:, :2] = g_mean2D
    return {"means3D": g_means3D, "means2D": m2, "shs": g_shs, "opacities": g_opac.reshape(n, 1),
            "scales": g_scales, "rots": g_rots, "conic": g_conic, "rgb": g_rgb}
