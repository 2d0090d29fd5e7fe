"""cProfile of the Python side of render() (no-grad, C3): where does the host spend its ~0.7 ms per view?"""
import cProfile
import importlib
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
g4d = importlib.import_module("4dgaussians_b200")
synth = importlib.import_module("4dgaussians_b200.synth")


class Pipe:
    convert_SHs_python = False
    compute_cov3D_python = False
    debug = False


w = synth.WORKLOADS["C3"]
scene = synth.make_scene(w["n"], seed=0, scale_mean=w["scale_mean"])
mod = g4d.deform_network(synth.hidden_args(w["net"]))
synth.perturb_deformation(mod, 0)
mod.deformation_net.set_aabb(scene["aabb"][0].tolist(), scene["aabb"][1].tolist())
mod = mod.cuda()
pc = synth.SyntheticGaussianModel(scene, mod, sh_degree=3, requires_grad=False)
cams = synth.orbit_cameras(64, w["width"], w["height"], radius=w["radius"], focal=w["focal"], timestamps=300)
bg = torch.tensor(w["bg"], dtype=torch.float32, device="cuda")
with torch.no_grad():
    for i in range(20):
        g4d.render(cams[i % 64], pc, Pipe, bg)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(200):
        g4d.render(cams[i % 64], pc, Pipe, bg)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("enqueue %.3f ms/view, total %.3f ms/view" % ((t1 - t0) / 200 * 1e3, (t2 - t0) / 200 * 1e3))
    pr = cProfile.Profile()
    pr.enable()
    for i in range(200):
        g4d.render(cams[i % 64], pc, Pipe, bg)
    pr.disable()
    torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
