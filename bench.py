#!/usr/bin/env python
"""bench.py -- headline benchmark of the g4d hot path (BASELINE.json: render FPS & train-step ms @300k Gaussians,
1352x1014, HBM GB/s vs roofline), one JSON line on stdout.

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA path through the public API)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU implementation of the path

A "step" is one pass of the hot path over one synthetic view: the fused deform + rasterize FORWARD of workload C3
(`configs[3]`, the configuration the metric is quoted on; 300k Gaussians fit one GPU).  Views / timestamps are sharded
over ranks with no data-path collective (weak scaling); the per-step NCCL gradient all-reduce only exists
in the training step, reported in `train_step` (B=2 views fwd+bwd + flat-bucket all-reduce + fused Adam).

Timing: CUDA events around every step on the launching stream, W warm-up steps, an L2 flush (write of a 512 MiB buffer)
between timed steps outside the event pairs, barrier + synchronize on both sides of the timed region, MAX over ranks.
`value` = inputs resident in HBM; `e2e` = the same steps through the public `render()` with the camera arriving from host
memory and the rendered image copied to pinned host memory every step (reference render.py:59-60 semantics).
"""
from __future__ import annotations

import argparse
import importlib
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

METRIC = "render FPS @300k gauss 1352x1014 (fused deform+rasterize forward); train-step ms in `train_step`"
UNIT = "frames/s"
WORKLOAD = "C3"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            j = json.load(open(p))
            return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def _tensor_roofline(n, geom_ms):
    """The geometry kernel is bound by the tensor pipe + its epilogues, not by HBM: report it against the measured dense
    BF16 peak too.  Algorithmic FLOPs = 2 * (F*Wd + h*Wd^2 + Wd*sum(k)) per Gaussian = 187.1 kFLOP (dynerf net)."""
    try:
        j = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak, src = float(j["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
    except Exception:
        peak, src = 2250.0, "fallback (nominal dense bf16 2.25 PFLOP/s)"
    flops = 2.0 * (32 * 128 + 5 * 128 * 128 + 128 * (3 + 3 + 4 + 1 + 48)) * n
    ach = flops / (geom_ms * 1e-3) / 1e12 if geom_ms > 0 else 0.0
    return {"bound": "tensor", "kernel": "deform_features + deform_tc_kernel (stage `geom`)", "achieved": ach, "peak": peak,
            "unit": "TFLOP/s", "frac": ach / peak if peak else None, "peak_source": src, "algorithmic_flops": flops,
            "kernel_ms": geom_ms,
            "note": "fp32-accurate 3xTF32: every algorithmic MAC is 3 tensor-core MACs at the TF32 rate (half of BF16), so "
                    "the ceiling of this scheme is peak/6; ncu: tensor pipe active 36 % of the kernel (profiles/r1e_ncu_full_C3.md)"}


class ClockSampler:
    """SM clock / throttle reasons DURING the timed region.  In-process NVML (three light queries every 200 ms from a thread;
    ctypes drops the GIL during the calls); falls back to a low-rate `nvidia-smi -lms` subprocess when pynvml is unusable."""
    Q = "index,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index: int):
        self.index, self.sm, self.mx, self.reasons, self.proc, self.nvml, self._stop = index, [], [], set(), None, None, False
        self.source = None

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        try:
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
        return pynvml, h

    def start(self):
        try:
            self.nvml, self.handle = self._nvml_handle()
            self.nvml.nvmlDeviceGetClockInfo(self.handle, self.nvml.NVML_CLOCK_SM)
            self.source = "nvml"
            threading.Thread(target=self._poll_nvml, daemon=True).start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi -lms 100"
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _poll_nvml(self):
        n = self.nvml
        while not self._stop:
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                self.mx.append(float(n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)))
                try:
                    r = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                except Exception:
                    r = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                for name, bit in self.BITS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.2)       # NVML queries occasionally stall the driver for milliseconds: keep them rare

    def _pump(self):
        for line in self.proc.stdout:
            c = [x.strip() for x in line.split(",")]
            if len(c) >= 7 and c[1].replace(".", "").isdigit():
                self.sm.append(float(c[1])); self.mx.append(float(c[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[3:7]):
                    if v.lower().startswith("active"):
                        self.reasons.add(name)

    def mark(self):
        """start of the timed region: earlier samples (NVML warm-up, GPU idle) are dropped"""
        self.sm, self.mx, self.reasons = [], [], set()

    def stop(self):
        self._stop = True
        if self.proc is not None:
            self.proc.terminate()
        if self.source is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock source (pynvml / nvidia-smi unavailable or disabled)"]}
        sm_sorted = sorted(self.sm)
        load = sm_sorted[len(sm_sorted) // 2:] if sm_sorted else []      # under load = upper half of the samples
        return {"sm_mhz": float(np.median(load)) if load else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "source": self.source}


def build_scene(device, seed=0):
    g4d = importlib.import_module("4dgaussians_b200")
    synth = importlib.import_module("4dgaussians_b200.synth")
    w = synth.WORKLOADS[WORKLOAD]
    scene = synth.make_scene(w["n"], seed=seed, scale_mean=w["scale_mean"])
    torch.manual_seed(seed)
    mod = g4d.deform_network(synth.hidden_args(w["net"]))
    synth.perturb_deformation(mod, seed)
    mod.deformation_net.set_aabb(scene["aabb"][0].tolist(), scene["aabb"][1].tolist())
    return g4d, synth, w, scene, mod


def algorithmic_bytes(n, H, W, R, param_bytes):
    """SURVEY §8d B_fwd per view and its per-stage split."""
    geom = n * 236 + n * 4 + param_bytes           # SoA read once + radii + planes/MLP read once
    binning = R * (8 + 4) * 2                      # one (key,id) record written then read
    blend = R * 44 + H * W * 16                    # per-instance fetch + colour/depth write
    return {"geom": geom, "binning": binning, "blend": blend, "total": geom + binning + blend}


class Pipe:
    convert_SHs_python = False
    compute_cov3D_python = False
    debug = False


def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the g4d path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        # keep stdout to the one JSON line: NCCL prints its version banner to stdout at levels VERSION and WARN
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            os.environ.pop("NCCL_DEBUG")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    g4d, synth, w, scene, mod = build_scene(dev)
    lib = g4d._lib
    mod = mod.to(dev)
    pc = synth.SyntheticGaussianModel(scene, mod, device=dev, sh_degree=3, requires_grad=False)
    K, Wm = args.steps, args.warmup
    n_views = (K + Wm) * world
    cams_all = synth.orbit_cameras(max(n_views, 8), w["width"], w["height"], radius=w["radius"], focal=w["focal"], timestamps=300)
    my_cams = [cams_all[(i * world + rank) % len(cams_all)] for i in range(K + Wm)]
    bg = torch.tensor(w["bg"], dtype=torch.float32, device=dev)
    ws = lib.Workspace.get(local)
    ws.set_option(lib.OPT_STAGE_TIMING, 1)
    ws.set_option(lib.OPT_SYNC_MODE, 0 if args.no_host_sync else 1)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    H, Wd = w["height"], w["width"]
    rz = importlib.import_module("4dgaussians_b200.renderer")

    def barrier():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ------------------------------------------------------------------ resident arm (`value`)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    stage_acc, Rs, vis = {}, [], []
    sampler = ClockSampler(local)
    if rank == 0 and not os.environ.get("G4D_BENCH_NO_SAMPLER"):
        sampler.start()          # NVML initialisation / first queries happen during the warm-up, not in the timed region
    import gc
    with torch.no_grad():
        for i in range(Wm):
            g4d.render(my_cams[i], pc, Pipe, bg)
        gc.collect(); gc.disable()     # a host hiccup shows up 1:1 in a ~1 ms step that synchronises on R once per frame
        barrier()
        # Ranks leave the collective barrier at different times and an idle B200 drops its clocks within milliseconds (measured:
        # the 2nd step after an idle gap stalls 5-40 ms).  Two more untimed warm-up renders, then the synchronize that
        # brackets the timed region -- after it the GPU is idle for microseconds only.
        for i in range(2):
            g4d.render(my_cams[i], pc, Pipe, bg)
        torch.cuda.synchronize(dev)
        sampler.mark()
        t_wall0 = time.perf_counter()
        for i in range(K):
            flush.fill_(i & 0xFF)
            ev[i][0].record()
            out = g4d.render(my_cams[Wm + i], pc, Pipe, bg)
            ev[i][1].record()
        barrier()
        t_wall = time.perf_counter() - t_wall0
        gc.enable()
        # per-stage device times and instance counts: separate, untimed pass (the queries synchronise)
        for i in range(min(K, 8)):
            g4d.render(my_cams[Wm + i], pc, Pipe, bg)
            torch.cuda.synchronize(dev)
            fn_ctx = ws._free_contexts[-1] if ws._free_contexts else None
            if fn_ctx is not None:
                for k_, v_ in fn_ctx.stage_times().items():
                    stage_acc[k_] = stage_acc.get(k_, 0.0) + v_
                s_ = fn_ctx.stats()
                Rs.append(int(s_.num_rendered)); vis.append(int(s_.num_visible))
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([sum(step_ms)], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    value = world * K / (total_ms / 1e3)

    # ------------------------------------------------------------------ e2e arm: host camera in, image out to pinned host
    pinned = torch.empty(3, H, Wd, dtype=torch.float32).pin_memory()
    ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    with torch.no_grad():
        barrier()
        for i in range(K):
            flush.fill_(i & 0xFF)
            ev2[i][0].record()
            out = g4d.render(my_cams[Wm + i], pc, Pipe, bg)          # camera matrices are HOST tensors -> kernel params
            pinned.copy_(out["render"], non_blocking=True)
            ev2[i][1].record()
        barrier()
    e2e_ms = torch.tensor([sum(a.elapsed_time(b) for a, b in ev2)], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = world * K / (float(e2e_ms.item()) / 1e3)
    cam_bytes = 4 * (16 + 16 + 3 + 3 + 4) + 16

    # ------------------------------------------------------------------ opt-in exact-image tile culling (same pixels, fewer bins)
    ws.set_option(lib.OPT_TIGHT_CULL, 1)
    ev3 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    with torch.no_grad():
        for i in range(3):
            g4d.render(my_cams[i], pc, Pipe, bg)
        barrier()
        for i in range(K):
            flush.fill_(i & 0xFF)
            ev3[i][0].record()
            out = g4d.render(my_cams[Wm + i], pc, Pipe, bg)
            ev3[i][1].record()
        barrier()
        fn_ctx = ws._free_contexts[-1] if ws._free_contexts else None
        R_tight = int(fn_ctx.stats().num_rendered) if fn_ctx is not None else None
    tight_ms = torch.tensor([sum(a.elapsed_time(b) for a, b in ev3)], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(tight_ms, op=dist.ReduceOp.MAX)
    tight_value = world * K / (float(tight_ms.item()) / 1e3)
    train_tight = run_train_steps(g4d, synth, lib, w, scene, mod, dev, dist, world, rank, max(3, min(K, 10)), 5, flush)
    ws.set_option(lib.OPT_TIGHT_CULL, 0)

    # ------------------------------------------------------------------ training step (B=2 views, fwd+bwd, all-reduce, Adam)
    train = run_train_steps(g4d, synth, lib, w, scene, mod, dev, dist, world, rank, max(3, min(K, 10)), 5, flush)
    clocks = sampler.stop() if rank == 0 else None     # sampled from the start of the timed region to the end of the train steps

    if rank == 0:
        peak, peak_src = _peaks()
        R = float(np.mean(Rs)) if Rs else 0.0
        pbytes = sum(p.numel() * 4 for p in mod.flat_parameters())
        ab = algorithmic_bytes(w["n"], H, Wd, R, pbytes)
        stages = {k_: v_ / max(1, len(Rs)) for k_, v_ in stage_acc.items()}
        fwd_stages = {k_: stages.get(k_, 0.0) for k_ in ("prep", "geom", "scan", "emit", "sort", "ranges", "blend")}
        dom = max(("geom", "blend", "sort"), key=lambda k_: fwd_stages.get(k_, 0.0))
        dom_bytes = {"geom": ab["geom"], "blend": ab["blend"], "sort": ab["binning"]}[dom]
        dom_ms = fwd_stages[dom]
        achieved = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
        traffic, traffic_src = None, None
        try:     # dram bytes per launch of the dominant kernel from the committed ncu --set full capture of this workload
            tj = json.load(open(os.path.join(ROOT, "profiles", "r1e_traffic.json")))
            bpl = tj["bytes_per_launch"]
            keys = {"geom": ["deform_tc_kernel<1, 16, 2, 1>", "deform_features_kernel<4>"], "blend": ["blend_forward_kernel<2>"], "sort": []}[dom]
            if keys and all(k_ in bpl for k_ in keys):
                traffic, traffic_src = float(sum(bpl[k_] for k_ in keys)), tj["source"]
        except Exception:
            pass
        ms_per_step = total_ms / K
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C3: 300k Gaussians, 1352x1014, dynerf net (C=16,L=2,T=150,Wd=128,5 heads), sh_degree 3, "
                                   "300 timestamps, views sharded over ranks", "n_gaussians": w["n"], "width": Wd, "height": H,
                       "focal_px": w["focal"], "orbit_radius": w["radius"], "scale_mean": w["scale_mean"],
                       "tile_instances_R": R, "visible_gaussians": float(np.mean(vis)) if vis else None,
                       "l2": "flushed between timed steps (512 MiB write, outside the event pairs)",
                       "binning": "capacity-bounded, no host sync (overflow-checked)" if args.no_host_sync else "host sync on R",
                       "mlp": "tcgen05 3xTF32 forward (fp32-accurate), BF16x2 tcgen05 backward",
                       "parallelism": "scene replicated, views sharded (dp%d)" % world},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": cam_bytes, "d2h_bytes_per_step": 3 * H * Wd * 4,
                    "how": "public render() per frame, camera from host memory, image copied to pinned host memory on the render "
                           "stream inside every step's event pair"},
            "gpu_launches": 8 * K,
            "gpu_launches_note": "own kernels per step: pack_camera, collapse_time_rows, deform_features, deform_tc_kernel(fused), "
                                 "depth_keys, emit_keys, tile_ranges, blend_forward; plus CUB scan/sort library launches",
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": {"geom": "deform_tc_kernel<1,16,2> (fused deform+activate+project, tcgen05)",
                                                    "blend": "blend_forward_kernel", "sort": "cub::DeviceRadixSort"}[dom],
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if peak else None,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "algorithmic_bytes": dom_bytes, "kernel_ms": dom_ms},
            "roofline_path": {"what": "whole fused forward, B_fwd of SURVEY 8d", "algorithmic_bytes": ab["total"],
                              "achieved": ab["total"] / (ms_per_step * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                              "frac": ab["total"] / (ms_per_step * 1e-3) / 1e9 / peak},
            "roofline_tensor": _tensor_roofline(w["n"], fwd_stages.get("geom", 0.0)),
            "stage_ms": stages, "train_step": train, "wall_s_timed_region": t_wall,
            "step_ms": [round(x, 3) for x in step_ms],
            "exact_tile_cull": {"what": "opt-in G4D_OPT_TIGHT_CULL: (Gaussian, tile) pairs whose best-case alpha over the tile is "
                                        "< 1/255 are not binned; pixels bit-identical (tests), bins no longer the reference's",
                                "value": tight_value, "unit": UNIT, "tile_instances_R_last_view": R_tight,
                                "train_step_ms": train_tight["ms_per_step"]},
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_reference_arm(steps=1, warmup=1, quiet=True)["cpu_baseline"]
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def run_train_steps(g4d, synth, lib, w, scene, mod, dev, dist, world, rank, steps, warmup, flush):
    """Full train step per rank: B=2 views (cook_spinach batch size) fused fwd + bwd, L1 loss against a resident target,
    ONE flat-bucket NCCL all-reduce of all gradients, fused Adam.  Returns ms per step (max over ranks)."""
    dp = importlib.import_module("4dgaussians_b200.dp")
    pc = synth.SyntheticGaussianModel(scene, mod, device=dev, sh_degree=3, requires_grad=True)
    params = pc.gaussian_parameters() + [p for p in mod.flat_parameters()]
    bucket = dp.FlatGradBucket(params)
    mod.fused_grad_accumulation = True      # backward kernels add straight into the bucket views (deformation.py)
    opt = torch.optim.Adam([{"params": params, "lr": 1e-4}], eps=1e-15, fused=True)
    B = 2
    cams = synth.orbit_cameras(64, w["width"], w["height"], radius=w["radius"], focal=w["focal"], timestamps=300)
    bg = torch.tensor(w["bg"], dtype=torch.float32, device=dev)
    target = torch.rand(3, w["height"], w["width"], device=dev)
    evs = []
    for it in range(steps + warmup):
        flush.fill_(it & 0xFF)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if it == warmup:
            torch.cuda.synchronize(dev)
            if dist is not None:
                dist.barrier()
        a.record()
        bucket.zero_()
        for v in range(B):
            cam = cams[((it * B + v) * world + rank) % len(cams)]
            out = g4d.render(cam, pc, Pipe, bg)
            loss = torch.nn.functional.l1_loss(out["render"], target) / B       # = utils/loss_utils.py:l1_loss
            loss.backward()
        bucket.allreduce_mean(dist, world)
        opt.step()
        b.record()
        if it >= warmup:
            evs.append((a, b))
    torch.cuda.synchronize(dev)
    ms = torch.tensor([sum(x.elapsed_time(y) for x, y in evs) / len(evs)], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ctx = lib.Workspace.get(dev.index)._free_contexts
    st = ctx[-1].stage_times() if ctx else {}
    return {"ms_per_step": float(ms.item()), "step_ms": [round(x.elapsed_time(y), 3) for x, y in evs],
            "views_per_step_per_gpu": B, "global_batch": B * world,
            "includes": "2x fused fwd+bwd (network gradients accumulated straight into the flat bucket), L1 loss, one flat-bucket all-reduce (%d floats), fused Adam" % bucket.numel,
            "last_view_stage_ms": st}


# ------------------------------------------------------------------------------------------------------------------
def cpu_reference_arm(steps: int, warmup: int, quiet: bool = False):
    """The reference's CPU implementation of the path on the host cores: deformation = the reference's own PyTorch module
    when /root/reference is present (build container), else the oracle port; rasterizer = oracle C port with OpenMP
    (the reference rasterizer's source is absent -> kind 'port').  One step = one full C3 view."""
    from oracle import deform_ref as dr
    from oracle import raster_ref as rr
    from oracle.ref_loader import reference_available, load_reference_deform_network
    g4d, synth, w, scene, mod = build_scene("cpu")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from util_scene import cam_tuple, oracle_params_from_module
    cores = os.cpu_count() or 1
    torch.set_num_threads(min(cores, 32))     # the element-wise CPU kernels stop scaling (and start thrashing) beyond that
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    cfg, prm = oracle_params_from_module(mod)
    for t_ in prm.leaves():
        t_.requires_grad_(False)
    kind = "port"
    ref_net = None
    if reference_available():
        ref_net = load_reference_deform_network(cfg)
        sd = ref_net.state_dict(); sd.update(dr.params_to_state_dict(prm)); ref_net.load_state_dict(sd)
        kind = "reference(deform)+port(rasterizer)"
    cams = synth.orbit_cameras(max(8, steps + warmup), w["width"], w["height"], radius=w["radius"], focal=w["focal"], timestamps=300)
    shs = torch.cat([scene["features_dc"], scene["features_rest"]], dim=1)
    n = w["n"]
    times, t_def, t_ras = [], [], []
    rr.lib()
    for i in range(steps + warmup):
        cam = cams[i % len(cams)]
        t0 = time.perf_counter()
        with torch.no_grad():
            if ref_net is not None:
                pts, sc, rot, op, sh = ref_net(scene["xyz"], scene["scaling"], scene["rotation"], scene["opacity"], shs,
                                               torch.tensor(cam.time).repeat(n, 1))
            else:
                pts, sc, rot, op, sh = dr.deform_forward(cfg, prm, scene["xyz"], scene["scaling"], scene["rotation"],
                                                         scene["opacity"], shs, cam.time)
            s, r, o = dr.activate(sc, rot, op)
        t1 = time.perf_counter()
        rc, _ = cam_tuple(cam, w["bg"], sh_degree=3)
        rr.rasterize_forward(rc, pts.numpy(), s.numpy(), r.numpy(), o.numpy(), sh.numpy())
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt); t_def.append(t1 - t0); t_ras.append(dt - (t1 - t0))
    ms = 1e3 * float(np.mean(times))
    fps = 1e3 / ms
    cb = {"value": fps, "unit": UNIT, "cores": cores, "kind": kind,
          "sample": "%d full C3 view(s) (300k Gaussians, 1352x1014): PyTorch-CPU deformation (%.2f s/view) + OpenMP C rasterizer "
                    "oracle (%.2f s/view), forward" % (len(times), float(np.mean(t_def)), float(np.mean(t_ras)))}
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": 1, "steps": steps, "warmup": warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": "C3: 300k Gaussians, 1352x1014, dynerf net, host CPU, %d threads" % cores},
            "cpu_baseline": cb, "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    return line


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = min(args.steps, 20)
    line = cpu_reference_arm(steps=steps, warmup=min(args.warmup, 3))
    line["n_gpus"] = args.gpus
    line["steps"] = steps
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="g4d", choices=["g4d", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-host-sync", dest="no_host_sync", action="store_true",
                    help="capacity-bounded binning without the per-forward host round trip (default: exact sizing with one "
                         "host sync per forward, the reference's behaviour -- measured faster at this scale because the "
                         "padded sort costs more than the bubble it removes)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
