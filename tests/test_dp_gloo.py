"""world_size-2 gloo test of the data-parallel host logic (FlatGradBucket, view sharding, stats reduction)."""
import importlib
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

dp = importlib.import_module("4dgaussians_b200.dp")


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    plane = torch.nn.Parameter(torch.rand(1, 4, 5, 3).contiguous(memory_format=torch.channels_last))
    w = torch.nn.Parameter(torch.rand(6, 4))
    xyz = torch.nn.Parameter(torch.rand(7, 3))
    bucket = dp.FlatGradBucket([xyz, plane, w])
    assert plane.grad.shape == plane.shape and plane.grad.stride() == plane.stride()
    views = dp.shard_views(6, rank, world)
    bucket.zero_()
    for v in views:      # "loss" of view v: (v+1) * sum of every parameter, mean over the global batch
        loss = (v + 1) * (xyz.sum() + 2 * plane.sum() + 3 * w.sum()) / len(views)
        loss.backward()
    flat_ptr = bucket.flat.data_ptr()
    bucket.allreduce_mean(dist, world)
    assert bucket.flat.data_ptr() == flat_ptr and xyz.grad.data_ptr() == flat_ptr     # grads are still views of the bucket
    gn = torch.full((7, 1), float(rank + 1)); den = torch.full((7, 1), 1.0); rad = torch.arange(7, dtype=torch.float32) * (rank + 1)
    dp.allreduce_densification_stats(dist, world, gn, den, rad)
    # plain numpy payloads: torch tensors travel through a queue as shared-memory handles that die with the sender
    q.put((rank, views) + tuple(x.detach().contiguous().numpy().copy() for x in (xyz.grad, plane.grad, w.grad, gn, den, rad)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_flat_bucket_allreduce_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=90) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    res = [t[:2] + tuple(torch.from_numpy(x) for x in t[2:]) for t in res]
    (r0, v0, gx0, gp0, gw0, gn0, den0, rad0), (r1, v1, gx1, gp1, gw1, gn1, den1, rad1) = res
    assert v0 == [0, 2, 4] and v1 == [1, 3, 5]
    # mean over the 6 views of (v+1) = 3.5
    assert torch.allclose(gx0, torch.full((7, 3), 3.5)) and torch.equal(gx0, gx1)
    assert torch.allclose(gp0, torch.full((1, 4, 5, 3), 7.0)) and torch.allclose(gw1, torch.full((6, 4), 10.5))
    assert torch.allclose(gn0, torch.full((7, 1), 3.0)) and torch.allclose(den1, torch.full((7, 1), 2.0))
    assert torch.equal(rad0, torch.arange(7, dtype=torch.float32) * 2) and torch.equal(rad0, rad1)


def test_single_rank_is_identity():
    p = torch.nn.Parameter(torch.ones(3, 2))
    b = dp.FlatGradBucket([p])
    (p * 2).sum().backward()
    b.allreduce_mean(None, 1)
    assert torch.equal(p.grad, torch.full((3, 2), 2.0)) and dp.shard_views(5, 0, 1) == [0, 1, 2, 3, 4]


def test_bucket_detects_detached_grads():
    """zero_grad(set_to_none=True) (train.py:292) detaches .grad from the bucket; the collective must not run over
    a stale buffer silently."""
    p = torch.nn.Parameter(torch.ones(3, 2)); q = torch.nn.Parameter(torch.ones(4))
    b = dp.FlatGradBucket([p, q])
    opt = torch.optim.SGD([p, q], lr=0.1)
    (p.sum() + q.sum()).backward()
    b.allreduce_mean(None, 1)
    opt.zero_grad(set_to_none=True)
    assert b.detached() == [0, 1]
    (3 * p.sum() + q.sum()).backward()            # autograd allocates fresh .grad tensors outside the bucket
    with pytest.raises(RuntimeError, match="no longer a view of the bucket"):
        b.allreduce_mean(None, 1)
    b.allreduce_mean(None, 1, on_detached="adopt")
    assert b.detached() == [] and torch.equal(b.flat, torch.cat([torch.full((6,), 3.0), torch.ones(4)]))
    b.zero_(); b.attach()
    (p.sum() * 2).backward()
    assert torch.equal(p.grad, torch.full((3, 2), 2.0)) and p.grad.data_ptr() == b.flat.data_ptr()


def _trainer_worker(rank, world, port, q):
    """DPTrainer's own exchange (train_dp.py): the flat gradient buffer in `comm_chunks` asynchronous slices and the
    densification statistics reduced on demand -- on CPU tensors over gloo (no kernels are launched)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    td = importlib.import_module("4dgaussians_b200.train_dp")
    synth = importlib.import_module("4dgaussians_b200.synth")
    g4d = importlib.import_module("4dgaussians_b200")
    torch.manual_seed(3)
    scene = synth.make_scene(300, seed=1, scale_mean=0.05)
    mod = g4d.deform_network(synth.hidden_args("small128"))
    gs = td.GaussianSet(scene, mod, device="cpu")
    tr = td.DPTrainer(gs, td.default_opt(), dist=dist, world_size=world, rank=rank, cameras_extent=2.0, seed=0)
    st = tr.state
    base = torch.arange(st.numel, dtype=torch.float32) % 97.0
    st.grad.copy_(base * (rank + 1))                      # rank r holds (r + 1) * base: the SUM over two ranks is 3 * base
    works = tr._launch_allreduce()
    assert len(works) == tr.comm_chunks and works[0][0] == 0 and works[-1][1] == st.numel
    assert all(b % 4 == 0 for b, _, _ in works)           # slices start on 16-byte boundaries (Adam kernel spans)
    for _, _, w in works:
        tr._wait(w)
    ok_grad = bool(torch.equal(st.grad, 3.0 * base)) and st.attached() and not tr._host_staged_collectives()
    # a CUDA buffer over gloo is handed over finished and awaited on the host (stand-in device, counting synchronisation)
    real_device, calls = st.device, []
    st.device = torch.device("cuda", 0)
    tr._sync_device = lambda: calls.append(1)

    class _Done:
        def wait(self):
            calls.append(0)
    ok_grad = ok_grad and tr._host_staged_collectives()
    tr._wait(_Done())
    st.device = real_device
    ok_grad = ok_grad and calls == [0, 1]
    n = gs._xyz.shape[0]
    gs.xyz_gradient_accum.fill_(float(rank + 1)); gs.denom.fill_(1.0)
    gs.max_radii2D.copy_(torch.arange(n, dtype=torch.float32) * (rank + 1))
    accum, denom, radii = tr._global_stats()
    ok_stats = bool(torch.equal(accum, torch.full((n, 1), 3.0)) and torch.equal(denom, torch.full((n, 1), 2.0))
                    and torch.equal(radii, torch.arange(n, dtype=torch.float32) * 2)
                    and float(gs.xyz_gradient_accum[0]) == rank + 1)          # the local accumulators stay local
    q.put((rank, ok_grad, ok_stats))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_trainer_sliced_allreduce_and_global_stats_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_trainer_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=150) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    assert res == [(0, True, True), (1, True, True)]
