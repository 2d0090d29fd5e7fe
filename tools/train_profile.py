"""Where does a train step go?  torch.profiler (kineto/CUPTI) over a few C3 train steps: kernel table + host table.
Never a bench number (it runs under a profiler)."""
import argparse
import importlib
import os
import sys
import time

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
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--profile", type=int, default=1)
    ap.add_argument("--flush", type=int, default=0)
    ap.add_argument("--stage-timing", type=int, default=0)
    ap.add_argument("--events", type=int, default=0)
    ap.add_argument("--smi-ms", type=int, default=0, help="run nvidia-smi -lms N alongside (sampler interference test)")
    a = ap.parse_args()
    g4d = importlib.import_module("4dgaussians_b200")
    synth = importlib.import_module("4dgaussians_b200.synth")
    dp = importlib.import_module("4dgaussians_b200.dp")
    w = synth.WORKLOADS[a.workload]
    dev = torch.device("cuda", 0)
    scene = synth.make_scene(w["n"], seed=0, scale_mean=w["scale_mean"])
    mod = g4d.deform_network(synth.hidden_args(w["net"]))
    synth.perturb_deformation(mod, 0)
    mod.deformation_net.set_aabb(scene["aabb"][0].tolist(), scene["aabb"][1].tolist())
    mod = mod.cuda()
    pc = synth.SyntheticGaussianModel(scene, mod, device=dev, sh_degree=3, requires_grad=True)
    params = pc.gaussian_parameters() + [p for p in mod.flat_parameters()]
    bucket = dp.FlatGradBucket(params)
    opt = torch.optim.Adam([{"params": params, "lr": 1e-4}], eps=1e-15, fused=True)
    cams = synth.orbit_cameras(64, w["width"], w["height"], radius=w["radius"], focal=w["focal"], timestamps=300)
    bg = torch.tensor(w["bg"], dtype=torch.float32, device=dev)
    target = torch.rand(3, w["height"], w["width"], device=dev)
    B = 2
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev) if a.flush else None
    if a.stage_timing:
        g4d._lib.Workspace.get(0).set_option(g4d._lib.OPT_STAGE_TIMING, 1)

    def step(it, marks=None):
        def mark(name):
            if marks is not None:
                torch.cuda.synchronize()
                marks.append((name, time.perf_counter()))
        mark("start")
        bucket.zero_()
        mark("zero")
        for v in range(B):
            cam = cams[(it * B + v) % len(cams)]
            out = g4d.render(cam, pc, Pipe, bg)
            mark("fwd%d" % v)
            loss = (out["render"] - target).abs().mean() / B
            mark("loss%d" % v)
            loss.backward()
            mark("bwd%d" % v)
        bucket.allreduce_mean(None, 1)
        opt.step()
        mark("adam")

    for it in range(3):
        step(it)
    torch.cuda.synchronize()
    smi = None
    if a.smi_ms:
        import subprocess
        smi = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,clocks_event_reasons.active", "--format=csv,noheader",
                                "-lms", str(a.smi_ms)], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        time.sleep(0.5)
    # (1) free-running wall time per step
    t0 = time.perf_counter()
    evs = []
    for it in range(a.steps):
        if flush is not None:
            flush.fill_(it & 0xFF)
        if a.events:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        step(3 + it)
        if a.events:
            e1.record(); evs.append((e0, e1))
    torch.cuda.synchronize()
    print("free-running: %.3f ms / step" % ((time.perf_counter() - t0) / a.steps * 1e3))
    if evs:
        print("event-timed:", " ".join("%.2f" % x.elapsed_time(y) for x, y in evs))
    # (2) serialised phases (sync after each): GPU+host time per phase
    for it in range(2):
        marks = []
        step(10 + it, marks)
        print("serialised:", " ".join("%s=%.2f" % (marks[i][0], (marks[i][1] - marks[i - 1][1]) * 1e3) for i in range(1, len(marks))),
              "total=%.2f" % ((marks[-1][1] - marks[0][1]) * 1e3))
    # (3) host-only time per phase (no syncs inside, host clock): how long does Python take to ENQUEUE a step
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    step(20)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("enqueue %.3f ms, drain %.3f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
    if smi is not None:
        smi.terminate()
    if a.profile:
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            for it in range(2):
                step(30 + it)
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=70))
        print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=25, max_name_column_width=70))


if __name__ == "__main__":
    main()
