"""Golden vectors for the pieces of the rasterizer math that DO exist in the reference tree
(the rasterizer itself is an absent submodule -- parity of the whole stage is unpinned):
  * utils/sh_utils.py:57-112 eval_sh           -> SH -> RGB polynomial
  * utils/general_utils.py:84-116 build_rotation / build_scaling_rotation -> Sigma = (RS)(RS)^T
  * scene/cameras.py:17-64 Camera + utils/graphics_utils.py -> view / projection matrices
TEST INFRASTRUCTURE ONLY.  Run in the build container:  python oracle/make_golden_raster_parts.py
"""
import os
import sys
import math

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle.ref_loader import _install_shims, REF_ROOT  # noqa: E402

_install_shims()
from utils.sh_utils import eval_sh  # type: ignore  # noqa: E402
from utils.graphics_utils import getWorld2View2, getProjectionMatrix, focal2fov, fov2focal  # type: ignore  # noqa: E402


def build_rotation_cpu(r):  # general_utils.build_rotation hard-codes device='cuda'; same formula on CPU
    import utils.general_utils as gu  # type: ignore
    src = open(os.path.join(REF_ROOT, "utils", "general_utils.py")).read()
    ns = {"torch": torch}
    start = src.index("def build_rotation(r):")
    end = src.index("def safe_state")
    exec(src[start:end].replace("device='cuda'", "device='cpu'").replace('device="cuda"', 'device="cpu"'), ns)
    return ns["build_rotation"](r), ns["build_scaling_rotation"]


def main():
    g = torch.Generator().manual_seed(42)
    n = 64
    dirs = torch.nn.functional.normalize(torch.randn(n, 3, generator=g, dtype=torch.float64), dim=-1)
    sh = torch.randn(n, 16, 3, generator=g, dtype=torch.float64)
    blob = {"dirs": dirs.numpy(), "sh": sh.numpy()}
    for deg in range(4):
        blob[f"rgb_deg{deg}"] = eval_sh(deg, sh.transpose(1, 2), dirs).numpy()
    q = torch.randn(n, 4, generator=g, dtype=torch.float64)
    s = torch.rand(n, 3, generator=g, dtype=torch.float64) * 0.3 + 0.01
    Rm, bsr = build_rotation_cpu(q.float())
    L = bsr(s.float(), q.float())
    blob["quat"] = q.numpy(); blob["scale"] = s.numpy()
    blob["rotmat_of_normalized_quat"] = Rm.numpy()
    blob["sigma"] = (L @ L.transpose(1, 2)).numpy()
    # cameras: D-NeRF style orbit exactly as dataset_readers.py:218-247 + cameras.py:59-64
    from scene.cameras import Camera  # type: ignore
    from importlib import import_module
    synth = import_module("4dgaussians_b200.synth")
    cams = []
    for theta, (w, h), fovx in ((-180.0, (400, 400), 0.6911112), (37.5, (1352, 1014), focal2fov(729.0, 1352)),
                                (121.0, (536, 960), focal2fov(480.0, 536))):
        c2w = synth.pose_spherical(theta, -30.0, 4.0)
        m = np.linalg.inv(np.array(c2w))
        R = -np.transpose(m[:3, :3]); R[:, 0] = -R[:, 0]; T = -m[:3, 3]
        fovy = focal2fov(fov2focal(fovx, w), h)
        cam = Camera(colmap_id=0, R=R, T=T, FoVx=fovx, FoVy=fovy, image=torch.zeros(3, h, w), gt_alpha_mask=None,
                     image_name="", uid=0, time=0.0)
        cams.append(dict(theta=theta, w=w, h=h, fovx=fovx, fovy=fovy, wvt=cam.world_view_transform.numpy(),
                         full=cam.full_proj_transform.numpy(), center=cam.camera_center.numpy()))
    for i, c in enumerate(cams):
        for k, v_ in c.items():
            blob[f"cam{i}_{k}"] = np.asarray(v_)
    out = os.path.join(os.path.dirname(HERE), "tests", "golden", "raster_parts.npz")
    np.savez_compressed(out, **blob)
    print("wrote", out, os.path.getsize(out) // 1024, "KiB")


if __name__ == "__main__":
    main()
