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
