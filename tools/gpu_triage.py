"""Stage-by-stage triage of the GPU path, each stage run by tools/gpu_triage.sh in its own process under a short timeout."""
import importlib
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
g4d = importlib.import_module("4dgaussians_b200")
synth = importlib.import_module("4dgaussians_b200.synth")
from util_scene import make_module, raster_inputs

stage = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2000


class Pipe:
    convert_SHs_python = False
    compute_cov3D_python = False
    debug = False


def settings(cam, bg):
    return g4d.GaussianRasterizationSettings(
        image_height=cam.image_height, image_width=cam.image_width, tanfovx=math.tan(cam.FoVx * 0.5),
        tanfovy=math.tan(cam.FoVy * 0.5), bg=torch.tensor(bg, dtype=torch.float32, device="cuda"), scale_modifier=1.0,
        viewmatrix=cam.world_view_transform.cuda(), projmatrix=cam.full_proj_transform.cuda(), sh_degree=3,
        campos=cam.camera_center.cuda(), prefiltered=False, debug=False)


import faulthandler
faulthandler.dump_traceback_later(35, exit=False)
print("stage", stage, "n", n, flush=True)
if stage == "raster":
    cam = synth.make_camera(10.0, 320, 240, radius=3.0)
    ins = [t.float().cuda() for t in raster_inputs(n, 1, scale_mean=0.03)]
    rast = g4d.GaussianRasterizer(settings(cam, (0, 0, 0)))
    with torch.no_grad():
        color, radii, depth = rast(means3D=ins[0], means2D=torch.zeros_like(ins[0]), shs=ins[4], colors_precomp=None,
                                   opacities=ins[3], scales=ins[1], rotations=ins[2], cov3D_precomp=None)
    torch.cuda.synchronize()
    print("raster ok", float(color.mean()), int((radii > 0).sum()), flush=True)
elif stage in ("deform128", "deform64"):
    mod = make_module("small128" if stage == "deform128" else "small64", seed=1)
    from oracle.make_golden_deform import synth_inputs
    (xyz, sc, rot, op, shs), _ = synth_inputs(n, 5)
    with torch.no_grad():
        outs = mod(xyz.cuda(), sc.cuda(), rot.cuda(), op.cuda(), shs.cuda(), torch.tensor(0.3).repeat(n, 1).cuda())
    torch.cuda.synchronize()
    print("deform ok", [float(o.float().mean()) for o in outs], flush=True)
elif stage == "fused":
    scene = synth.make_scene(n, seed=3, scale_mean=0.06)
    mod = make_module("small128", seed=1, aabb=scene["aabb"])
    pc = synth.SyntheticGaussianModel(scene, mod, sh_degree=3, requires_grad=False)
    cam = synth.make_camera(25.0, 128, 96, time=0.35)
    with torch.no_grad():
        out = g4d.render(cam, pc, Pipe, torch.tensor([0.1, 0.2, 0.3], device="cuda"))
    torch.cuda.synchronize()
    print("fused ok", float(out["render"].mean()), flush=True)
elif stage == "raster_bwd":
    cam = synth.make_camera(10.0, 320, 240, radius=3.0)
    ins = [t.float().cuda().requires_grad_(True) for t in raster_inputs(n, 1, scale_mean=0.03)]
    rast = g4d.GaussianRasterizer(settings(cam, (0, 0, 0)))
    m2 = torch.zeros_like(ins[0], requires_grad=True)
    color, radii, depth = rast(means3D=ins[0], means2D=m2, shs=ins[4], colors_precomp=None,
                               opacities=ins[3], scales=ins[1], rotations=ins[2], cov3D_precomp=None)
    torch.cuda.synchronize(); print("raster fwd ok", flush=True)
    color.sum().backward()
    torch.cuda.synchronize()
    print("raster bwd ok", float(ins[0].grad.abs().sum()), flush=True)
elif stage in ("deform_bwd128", "deform_bwd64"):
    mod = make_module("small128" if stage == "deform_bwd128" else "small64", seed=1)
    from oracle.make_golden_deform import synth_inputs
    (xyz, sc, rot, op, shs), _ = synth_inputs(n, 5)
    leaves = [t.cuda().requires_grad_(True) for t in (xyz, sc, rot, op, shs)]
    outs = mod(*leaves, torch.tensor(0.3).repeat(n, 1).cuda())
    torch.cuda.synchronize(); print("deform fwd(grad) ok", flush=True)
    sum(o.sum() for o in outs).backward()
    torch.cuda.synchronize()
    print("deform bwd ok", float(leaves[0].grad.abs().sum()), flush=True)
elif stage in ("fused_bwd", "fused_bwd_debug"):
    Pipe.debug = stage.endswith("debug")
    scene = synth.make_scene(n, seed=3, scale_mean=0.06)
    mod = make_module("small128", seed=1, aabb=scene["aabb"])
    pc = synth.SyntheticGaussianModel(scene, mod, sh_degree=3, requires_grad=True)
    cam = synth.make_camera(25.0, 128, 96, time=0.35)
    out = g4d.render(cam, pc, Pipe, torch.tensor([0.1, 0.2, 0.3], device="cuda"))
    torch.cuda.synchronize(); print("fused fwd(grad) ok", flush=True)
    out["render"].sum().backward()
    torch.cuda.synchronize()
    print("fused bwd ok", float(pc._xyz.grad.abs().sum()), flush=True)
elif stage == "smoke":
    import __graft_entry__ as ge
    ge.smoke()
