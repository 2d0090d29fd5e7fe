"""Tiny driver for ncu: a few fused fwd(+bwd) steps of a BASELINE workload through the public render() API.
    ncu ... python tools/profile_step.py --workload C3 --iters 2 --backward 1
Never a bench number (it runs under a profiler)."""
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
    ap.add_argument("--iters", type=int, default=2)
    ap.add_argument("--backward", type=int, default=1)
    ap.add_argument("--stage-times", type=int, default=0)
    ap.add_argument("--tc-debug", type=int, default=0)
    ap.add_argument("--tight-cull", type=int, default=0)
    ap.add_argument("--tc-mode", type=int, default=2, help="0 FFMA, 1 3xTF32, 2 FP16x2 two-slot (default)")
    a = ap.parse_args()
    g4d = importlib.import_module("4dgaussians_b200")
    synth = importlib.import_module("4dgaussians_b200.synth")
    w = synth.WORKLOADS[a.workload]
    scene = synth.make_scene(w["n"], seed=0, scale_mean=w["scale_mean"])
    mod = g4d.deform_network(synth.hidden_args(w["net"]))
    synth.perturb_deformation(mod, 0)
    mod.deformation_net.set_aabb(scene["aabb"][0].tolist(), scene["aabb"][1].tolist())
    mod = mod.cuda()
    pc = synth.SyntheticGaussianModel(scene, mod, sh_degree=3, requires_grad=bool(a.backward))
    cams = synth.orbit_cameras(8, w["width"], w["height"], radius=w["radius"], focal=w["focal"], timestamps=300)
    bg = torch.tensor(w["bg"], dtype=torch.float32, device="cuda")
    target = torch.rand(3, w["height"], w["width"], device="cuda")
    ws = g4d._lib.Workspace.get(0)
    if a.stage_times:
        ws.set_option(g4d._lib.OPT_STAGE_TIMING, 1)
    if a.tc_debug:
        ws.set_option(g4d._lib.OPT_TC_DEBUG, 1)
    if a.tight_cull:
        ws.set_option(g4d._lib.OPT_TIGHT_CULL, 1)
    ws.set_option(g4d._lib.OPT_TENSOR_CORES, a.tc_mode)
    for i in range(a.iters):
        cam = cams[i % len(cams)]
        if a.backward:
            out = g4d.render(cam, pc, Pipe, bg)
            (out["render"] - target).abs().mean().backward()
        else:
            with torch.no_grad():
                out = g4d.render(cam, pc, Pipe, bg)
        torch.cuda.synchronize()
        if a.tc_debug:
            import ctypes as C
            arr = (C.c_double * 12)()
            g4d._lib.load().g4d_debug_tc_cycles(ws.handle, arr)
            names = ["inputs", "wait_feat", "L0_mma", "wait_scratch_free", "epi0", "wait_W1", "L1_mma", "wait_W2", "epi_half",
                     "L2_mma", "head_out", "tail"]
            if a.tc_mode == 2:   # g4d_deform_f16.cu: thread 0 (slot 0) phases, then the MMA thread's blocked / total cycles
                names = ["E:wait_L0", "E:epi0", "E:wait_head", "E:small_head_epi", "E:sh_hidden_epi", "E:handover", "E:wait_L2+stage_feat",
                         "F:wait_deltas", "F:sh_out", "F:tail", "mma_warp_blocked", "mma_warp_total"]
            if a.backward:   # the backward kernel ran last and overwrote the slots (g4d_deform_tc_bwd.cu)
                ph = ["feat+L0", "epi0+dh", "dout", "head_epi", "wait_other", "mma+tail"]
                names = ["M:" + x for x in ph] + ["G:" + x for x in ph]
            print("tc cycles/CTA:", {k: int(arr[i]) for i, k in enumerate(names)}, flush=True)
        if a.stage_times and ws._free_contexts:
            c = ws._free_contexts[-1]
            s = c.stats()
            print("iter", i, "R=%d visible=%d" % (s.num_rendered, s.num_visible),
                  {k: round(v, 4) for k, v in c.stage_times().items()}, flush=True)
            ph = c.read("bin_phases")
            names = ["key_range", "depth_sort", "chunking", "count", "scan"]
            print("   bin_sort phase cycles (CTA 0):", {n_: int(ph[j + 1] - ph[j]) for j, n_ in enumerate(names)}, "key bits", int(ph[15]),
                  "| first radix pass:", {n_: int(ph[j + 1] - ph[j]) for j, n_ in zip(range(6, 10), ["hist->barrier", "barrier", "offsets", "scatter"])}
                  | {"hist": int(ph[6] - ph[1]), "barrier2": int(ph[10] - ph[9])}, flush=True)


if __name__ == "__main__":
    main()
