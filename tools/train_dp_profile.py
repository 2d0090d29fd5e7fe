"""Where does a DPTrainer.step go?  torch.profiler (kineto / CUPTI) over a few C3 train steps of the data-parallel harness
(4dgaussians_b200/train_dp.py): kernel table + the host-side top of the list.  Never a bench number (it runs under a profiler);
the un-profiled step time is printed first (CUDA events, 10 steps)."""
import argparse
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class Pipe:
    convert_SHs_python = False
    compute_cov3D_python = False
    debug = False


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="C3")
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    synth = importlib.import_module("4dgaussians_b200.synth")
    g4d = importlib.import_module("4dgaussians_b200")
    td = importlib.import_module("4dgaussians_b200.train_dp")
    w = synth.WORKLOADS[a.workload]
    dev = torch.device("cuda", 0)
    scene = synth.make_scene(w["n"], seed=0, scale_mean=w["scale_mean"])
    mod = g4d.deform_network(synth.hidden_args(w["net"]))
    synth.perturb_deformation(mod, 0)
    mod.deformation_net.set_aabb(scene["aabb"][0].tolist(), scene["aabb"][1].tolist())
    mod = mod.cuda()
    gs = td.GaussianSet(scene, mod, device=dev, sh_degree=3)
    tr = td.DPTrainer(gs, td.default_opt(), dist=None, world_size=1, rank=0, cameras_extent=2.6, seed=0)
    cams = synth.orbit_cameras(64, w["width"], w["height"], radius=w["radius"], focal=w["focal"], timestamps=300)
    bg = torch.tensor(w["bg"], dtype=torch.float32, device=dev)
    target = torch.rand(3, w["height"], w["width"], device=dev)
    B = 2

    def step(it):
        tr.step([cams[(it * B + v) % len(cams)] for v in range(B)], [target] * B, bg, Pipe)
    for it in range(5):
        step(it)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for it in range(5, 15):
        step(it)
    e1.record()
    torch.cuda.synchronize()
    print("un-profiled: %.3f ms / step (10 steps, CUDA events)" % (e0.elapsed_time(e1) / 10))
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for it in range(15, 15 + a.steps):
            step(it)
        torch.cuda.synchronize()
    print("=== kernels, %d steps (divide by %d) ===" % (a.steps, a.steps))
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
    print("=== host ===")
    print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=25, max_name_column_width=70))


if __name__ == "__main__":
    main()
