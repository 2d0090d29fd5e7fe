"""Make the reference's own import statements resolve to the g4d implementations.

    import importlib; importlib.import_module("4dgaussians_b200.dropin").install()
    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer   # g4d
    # optional (install(patch_reference=True) with /path/to/4DGaussians on sys.path):
    #   gaussian_renderer.render -> g4d fused render, scene.deformation.deform_network -> g4d module

See INTEGRATION.md for the two-line change a maintainer of the reference would make instead.
"""
from __future__ import annotations

import sys
import types


def install(patch_reference: bool = False):
    from . import deformation, rasterizer, renderer
    mod = types.ModuleType("diff_gaussian_rasterization")
    mod.GaussianRasterizationSettings = rasterizer.GaussianRasterizationSettings
    mod.GaussianRasterizer = rasterizer.GaussianRasterizer
    mod.__doc__ = "g4d drop-in for depth-diff-gaussian-rasterization (see 4dgaussians_b200/rasterizer.py)"
    sys.modules["diff_gaussian_rasterization"] = mod
    from . import simple_knn
    knn = types.ModuleType("simple_knn")
    knn_c = types.ModuleType("simple_knn._C")
    knn_c.distCUDA2 = simple_knn.distCUDA2
    knn._C = knn_c
    knn.__doc__ = "g4d drop-in for simple-knn (see 4dgaussians_b200/simple_knn.py)"
    sys.modules["simple_knn"], sys.modules["simple_knn._C"] = knn, knn_c
    if patch_reference:
        try:
            import scene.deformation as sd  # type: ignore
            sd.deform_network = deformation.deform_network
        except Exception:
            pass
        try:
            import gaussian_renderer as gr  # type: ignore
            gr.render = renderer.render
        except Exception:
            pass
    return mod
