"""N4: distCUDA2 replacement (mean squared distance to the 3 nearest neighbours) against the CPU oracle."""
import importlib

import numpy as np
import pytest
import torch

from oracle import knn_ref

pytestmark = pytest.mark.gpu
knn = importlib.import_module("4dgaussians_b200.simple_knn")


def _cloud(n, seed, kind):
    g = np.random.default_rng(seed)
    if kind == "uniform":
        return (g.random((n, 3), dtype=np.float32) * 2 - 1) * np.float32(1.3)
    if kind == "clustered":      # COLMAP-like: dense blobs + sparse background, anisotropic extent
        c = g.normal(size=(8, 3)) * [3.0, 1.0, 0.2]
        pts = c[g.integers(0, 8, n)] + g.normal(size=(n, 3)) * g.choice([0.01, 0.1, 0.5], size=(n, 1))
        return pts.astype(np.float32)
    if kind == "planar":         # degenerate extent along z, duplicates
        pts = g.random((n, 3), dtype=np.float32); pts[:, 2] = 0.25
        pts[: n // 10] = pts[n // 10: 2 * (n // 10)]
        return pts
    raise ValueError(kind)


@pytest.mark.parametrize("kind", ["uniform", "clustered", "planar"])
@pytest.mark.parametrize("n", [4, 5, 777, 6000])
def test_dist2_bit_exact_vs_bruteforce(n, kind):
    pts = _cloud(n, n, kind)
    got = knn.distCUDA2(torch.from_numpy(pts).cuda()).cpu().numpy()
    want = knn_ref.dist2_knn3_bruteforce(pts)
    assert got.shape == (n,) and np.array_equal(got, want), float(np.abs(got - want).max())


def test_dist2_large_vs_kdtree_and_dropin_name():
    importlib.import_module("4dgaussians_b200.dropin").install()
    from simple_knn._C import distCUDA2
    pts = _cloud(300_000, 1, "uniform")
    got = distCUDA2(torch.from_numpy(pts).cuda()).cpu().numpy()
    want = knn_ref.dist2_knn3_kdtree(pts)
    assert np.abs(got - want).max() <= 1e-6 * want.max()
    pts = _cloud(100_000, 2, "clustered")
    got = distCUDA2(torch.from_numpy(pts).cuda()).cpu().numpy()
    want = knn_ref.dist2_knn3_kdtree(pts)
    assert np.abs(got - want).max() <= 1e-5 * want.max()


def test_fewer_than_four_points_and_cpu_refusal():
    for n in (1, 2, 3):
        pts = _cloud(n, 9, "uniform")
        got = knn.distCUDA2(torch.from_numpy(pts).cuda()).cpu().numpy()
        assert np.all(got > 1e30)                       # fewer than 3 neighbours: the missing ones count as FLT_MAX
    assert knn.distCUDA2(torch.zeros(0, 3, device="cuda")).numel() == 0
    with pytest.raises(RuntimeError):
        knn.distCUDA2(torch.zeros(5, 3))
