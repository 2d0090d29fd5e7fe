"""N1 on the GPU: the flat Adam kernel against torch.optim.Adam, a short training run with densification / pruning through the
real render path, and replica consistency of two data-parallel ranks (gloo over CUDA tensors: one GPU is enough)."""
import importlib
import os
import socket

import numpy as np
import pytest
import torch

from util_scene import g4d, make_module, synth

pytestmark = pytest.mark.gpu
td = importlib.import_module("4dgaussians_b200.train_dp")


class _Pipe:
    convert_SHs_python = False
    compute_cov3D_python = False
    debug = False


def test_flat_adam_matches_torch_adam():
    torch.manual_seed(0)
    shapes = [(1001, 3), (77,), (5, 4, 3, 2), (64, 64)]
    ps = [torch.nn.Parameter(torch.randn(s, device="cuda")) for s in shapes]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    lrs = [1e-2, 3e-3, 5e-2, 1e-3]
    opt = torch.optim.Adam([{"params": [r], "lr": lr} for r, lr in zip(ref, lrs)], lr=0.0, eps=1e-15)
    st = td.FlatState([("g%d" % i, [p]) for i, p in enumerate(ps)], torch.device("cuda", 0))
    for it in range(5):
        st.zero_grad()
        for p, r in zip(ps, ref):
            gr = torch.randn_like(p) * (10.0 ** (it - 2))
            p.grad.copy_(gr * 4.0)                  # the kernel divides by the world size (grad_scale = 1/4)
            r.grad = gr.clone()
        st.adam_step({"g%d" % i: lr for i, lr in enumerate(lrs)}, grad_scale=0.25)
        opt.step()
        for p, r in zip(ps, ref):
            assert float((p.data - r.data).abs().max()) <= 2e-6 * max(1.0, float(r.data.abs().max())), it
    assert st.attached()


def _make_trainer(n=1500, seed=0, dist=None, world=1, rank=0):
    scene = synth.make_scene(n, seed=5, scale_mean=0.05)
    mod = make_module("small128", seed=2, aabb=scene["aabb"])
    gs = td.GaussianSet(scene, mod)
    opt = td.default_opt()
    opt.densify_from_iter, opt.pruning_from_iter, opt.min_gaussians_for_prune = 2, 2, 100
    opt.densification_interval = opt.pruning_interval = 4
    opt.densify_grad_threshold_fine_init = opt.densify_grad_threshold_after = 5e-6
    return td.DPTrainer(gs, opt, dist=dist, world_size=world, rank=rank, cameras_extent=2.6, seed=seed)


def test_training_steps_with_densification_reduce_the_loss():
    tr = _make_trainer()
    W, H = 96, 80
    cams = [synth.make_camera(th, W, H, radius=4.0, time=t) for th, t in ((0.0, 0.1), (60.0, 0.5), (120.0, 0.9), (200.0, 0.3))]
    bg = torch.zeros(3, device="cuda")
    gts = [torch.full((3, H, W), 0.25, device="cuda") for _ in cams]
    first, last, n0 = None, None, tr.g._xyz.shape[0]
    for it in range(12):
        loss = tr.step(cams[(it % 2) * 2:(it % 2) * 2 + 2], gts[:2], bg, _Pipe)
        if it == 0:
            first = float(loss)
        last = float(loss)
    assert tr.rebuilds > 1 and tr.g._xyz.shape[0] != n0          # N changed, the flat buffers were rebuilt
    assert tr.state.attached() and last < first, (first, last)
    assert bool(torch.isfinite(tr.state.param).all())


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _rank_main(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(0)
    tr = _make_trainer(dist=dist, world=world, rank=rank)
    W, H = 64, 48
    cams = [synth.make_camera(30.0 * i, W, H, radius=4.0, time=(i % 5) / 4.0) for i in range(12)]
    gts = [torch.full((3, H, W), 0.1 + 0.05 * (i % 4), device="cuda") for i in range(12)]
    bg = torch.zeros(3, device="cuda")
    for it in range(9):
        mine = [(it * 2 * world + v * world + rank) % 12 for v in range(2)]        # round-robin views of the global batch
        tr.step([cams[i] for i in mine], [gts[i] for i in mine], bg, _Pipe)
    torch.cuda.synchronize()
    q.put((rank, tr.g._xyz.shape[0], tr.rebuilds, tr.state.param.detach().cpu().numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_ranks_stay_bit_identical_through_densification():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=500) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (_, n0, rb0, p0), (_, n1, rb1, p1) = res
    assert n0 == n1 and rb0 == rb1 and rb0 > 1
    assert np.array_equal(p0, p1)          # replicas never diverge: same reduced gradients, same decisions, same noise
