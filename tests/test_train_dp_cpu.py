"""N1 host logic on CPU (no kernels): flat parameter / gradient layout, learning-rate schedule, rank-consistent densify /
prune with Adam moments carried like the reference's cat_tensors_to_optimizer / _prune_optimizer."""
import importlib
import math

import torch

td = importlib.import_module("4dgaussians_b200.train_dp")
synth = importlib.import_module("4dgaussians_b200.synth")
g4d = importlib.import_module("4dgaussians_b200")


def _trainer(seed=0, n=400):
    torch.manual_seed(3)
    scene = synth.make_scene(n, seed=1, scale_mean=0.05)
    mod = g4d.deform_network(synth.hidden_args("small128"))
    gs = td.GaussianSet(scene, mod, device="cpu")
    opt = td.default_opt()
    opt.densify_from_iter, opt.pruning_from_iter, opt.min_gaussians_for_prune = 0, 0, 10
    return td.DPTrainer(gs, opt, cameras_extent=2.0, seed=seed)


def test_flat_state_views_and_segments():
    tr = _trainer()
    st, g = tr.state, tr.g
    assert st.attached() and st.numel % 4 == 0
    names = [n for n, _ in st.groups]
    assert names == ["xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation", "grid", "deformation"]
    b, e = st.segments["xyz"]
    assert (b, e) == (0, 400 * 3) and st.segments["f_dc"] == (1200, 2400)
    assert all(off % 4 == 0 for _, _, _, off, _ in st._views)         # 16-byte aligned starts (vector loads in the kernels)
    g._xyz.data[5, 1] = 42.0
    assert float(st.param[5 * 3 + 1]) == 42.0                        # .data is a view of the flat buffer
    plane = g._deformation.deformation_net.grid.grids[0][0]
    assert plane.is_contiguous(memory_format=torch.channels_last) and plane.grad.stride() == plane.stride()
    (g._xyz.sum() * 2).backward()
    off, n = st.slices(g._xyz)
    assert torch.equal(st.grad[off:off + n], torch.full((n,), 2.0))
    g._xyz.grad = None
    assert not st.attached()


def test_expon_lr_matches_reference_formula():
    o = td.default_opt()
    f = lambda s: td.expon_lr(s, o.position_lr_init, o.position_lr_final, 0, o.position_lr_delay_mult, o.position_lr_max_steps)
    assert abs(f(0) - o.position_lr_init) < 1e-12 and abs(f(20_000) - o.position_lr_final) < 1e-12
    assert abs(f(10_000) - math.sqrt(o.position_lr_init * o.position_lr_final)) < 1e-12      # log-linear midpoint
    assert td.expon_lr(5, 0.0, 0.0) == 0.0


def test_densify_and_prune_are_deterministic_and_carry_moments():
    res = []
    for _ in range(2):                                  # two "ranks" with identical statistics and the same seed
        tr = _trainer(seed=7)
        g = tr.g
        n = g._xyz.shape[0]
        gen = torch.Generator().manual_seed(11)
        g.xyz_gradient_accum = torch.rand(n, 1, generator=gen) * 4e-4
        g.denom = torch.ones(n, 1)
        g._scaling.data[: n // 2] = math.log(0.1)      # big ones split ...
        g._scaling.data[n // 2:] = math.log(0.005)     # ... small ones clone
        m, v = tr.state.moments(g._xyz)
        m.copy_(torch.arange(n * 3, dtype=torch.float32).view(n, 3)); v.fill_(2.0)
        old_xyz, old_m = g._xyz.data.clone(), m.clone()
        grads = (g.xyz_gradient_accum / g.denom).squeeze(-1)
        small = torch.exp(g._scaling.data).max(dim=1).values <= 0.01 * 2.0
        n_clone, n_split = int(((grads >= 2e-4) & small).sum()), int(((grads >= 2e-4) & ~small).sum())
        assert n_clone > 0 and n_split > 0
        assert tr.densify(2e-4, iteration=600)
        n2 = g._xyz.shape[0]
        assert n2 == n + n_clone + n_split               # + clones, + 2 children - 1 parent per split
        keep = ~((grads >= 2e-4) & ~small)
        m2, v2 = tr.state.moments(g._xyz)
        assert torch.equal(g._xyz.data[: int(keep.sum())], old_xyz[keep]) and torch.equal(m2[: int(keep.sum())], old_m[keep])
        assert float(m2[int(keep.sum()):].abs().max()) == 0.0 and float(v2[int(keep.sum()):].abs().max()) == 0.0
        assert tr.state.attached() and g.denom.shape == (n2, 1) and float(g.denom.sum()) == 0.0
        # children are scaled down by 1 / (0.8 * 2) (gaussian_model.py:440)
        assert torch.allclose(g._scaling.data[-1], torch.log(torch.tensor(0.1 / 1.6)).expand(3))
        g._opacity.data[::3] = -10.0
        assert tr.prune_points((torch.sigmoid(g._opacity.data) < 0.005).squeeze(-1))
        assert g._xyz.shape[0] == n2 - len(range(0, n2, 3))
        tr.reset_opacity()
        assert float(torch.sigmoid(g._opacity.data).max()) <= 0.01 + 1e-6
        res.append({k: p.data.clone() for k, p in g.named().items()})
    for k in res[0]:
        assert torch.equal(res[0][k], res[1][k]), k       # identical decisions and identical split noise
