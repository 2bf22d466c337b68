"""GPU tests of the uniform-grid search (d <= 3, the reference's kd-tree regime): it must agree
bit for bit with the SIMT path (both rank on the same float64 d2 and (d2, index) key) and with the
oracle on the reference's 2-D configurations."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import mepol_b200
from mepol_b200 import _lib
from conftest import assert_knn_matches


def _data(kind, n, d, rng):
    if kind == "gauss":
        return rng.standard_normal((n, d))
    if kind == "uniform":
        return rng.uniform(-6, 6, (n, d))
    if kind == "traj":
        t = max(1, n // 400)
        x = np.cumsum(rng.standard_normal((t, (n + t - 1) // t, d)) * 0.05, axis=1).reshape(-1, d)
        return x[:n]
    if kind == "clustered":
        c = rng.uniform(-100, 100, (8, d))
        return c[rng.integers(0, 8, n)] + 1e-3 * rng.standard_normal((n, d))
    raise ValueError(kind)


@pytest.mark.parametrize("kind", ["gauss", "uniform", "traj", "clustered"])
@pytest.mark.parametrize("n,d,k1", [(8000, 2, 5), (24000, 2, 51), (5000, 3, 5), (3000, 1, 5), (700, 2, 60),
                                    (9000, 3, 33), (64, 2, 5), (5, 2, 5), (40000, 2, 5)])
def test_grid_and_simt_are_bit_identical(kind, n, d, k1):
    rng = np.random.default_rng(n * 7 + d + k1)
    x = torch.from_numpy(_data(kind, n, d, rng).astype(np.float32)).cuda()
    a = mepol_b200.knn_device(x, k1, algo=_lib.KNN_SIMT)
    b = mepol_b200.knn_device(x, k1, algo=_lib.KNN_GRID)
    assert torch.equal(a.indices, b.indices)
    assert torch.equal(a.distances, b.distances)
    assert torch.equal(a.rk, b.rk)


def test_grid_with_duplicates_and_exact_ties():
    rng = np.random.default_rng(3)
    base = rng.integers(-5, 6, (3000, 2)).astype(np.float32)          # lattice points: masses of exact ties
    x = torch.from_numpy(base).cuda()
    a = mepol_b200.knn_device(x, 5, algo=_lib.KNN_SIMT)
    b = mepol_b200.knn_device(x, 5, algo=_lib.KNN_GRID)
    assert torch.equal(a.indices, b.indices)
    assert torch.equal(a.distances, b.distances)
    same = torch.zeros((4000, 2), device="cuda")                      # every point identical: one cell, all ties
    r = mepol_b200.knn_device(same, 5, algo=_lib.KNN_GRID)
    assert torch.equal(r.indices.cpu(), torch.arange(5, dtype=r.indices.dtype).expand(4000, 5))
    assert float(r.distances.abs().max()) == 0.0


def test_grid_separate_queries_inside_and_outside_the_box():
    from oracle import knn_oracle
    rng = np.random.default_rng(9)
    x = rng.uniform(-1, 1, (6000, 2)).astype(np.float32)
    q = rng.uniform(-1.5, 1.5, (500, 2)).astype(np.float32)
    q[:5] = np.array([[50, 0], [-30, 40], [0, -1e3], [1.0, 1.0], [-1.0, -1.0]], dtype=np.float32)
    res = mepol_b200.knn_device(torch.from_numpy(x).cuda(), 5, queries=torch.from_numpy(q).cuda(), algo=_lib.KNN_GRID)
    ref = mepol_b200.knn_device(torch.from_numpy(x).cuda(), 5, queries=torch.from_numpy(q).cuda(), algo=_lib.KNN_SIMT)
    assert torch.equal(res.indices, ref.indices)
    assert torch.equal(res.distances, ref.distances)
    ref_d, ref_i = knn_oracle.knn_c(x, 5, queries=q, mode="kd_tree")
    assert_knn_matches(res.distances.cpu().numpy(), res.indices.cpu().numpy(), ref_d, ref_i, x, queries=q)


def test_grid_query_slice_of_the_fitted_set():
    rng = np.random.default_rng(10)
    x = torch.from_numpy(rng.standard_normal((9000, 2)).astype(np.float32)).cuda()
    full = mepol_b200.knn_device(x, 5, algo=_lib.KNN_GRID)
    part = mepol_b200.knn_device(x, 5, queries=x[3000:7000], algo=_lib.KNN_GRID)
    assert torch.equal(part.indices, full.indices[3000:7000])
    assert torch.equal(part.distances, full.distances[3000:7000])


def test_auto_picks_the_grid_for_tiny_dimensions():
    x = torch.randn((5000, 2), device="cuda")
    auto = mepol_b200.knn_device(x, 5)
    grid = mepol_b200.knn_device(x, 5, algo=_lib.KNN_GRID)
    assert torch.equal(auto.indices, grid.indices)
    with pytest.raises(_lib.MepolError):
        mepol_b200.knn_device(torch.randn((5000, 4), device="cuda"), 5, algo=_lib.KNN_GRID)


def test_grid_non_finite_rows_fall_back():
    x = torch.randn((3000, 2), device="cuda")
    q = torch.randn((100, 2), device="cuda")
    q[7, 0] = float("nan")
    res = mepol_b200.knn_device(x, 5, queries=q, algo=_lib.KNN_GRID)
    ok = torch.ones(100, dtype=torch.bool, device="cuda")
    ok[7] = False
    ref = mepol_b200.knn_device(x, 5, queries=q, algo=_lib.KNN_SIMT)
    assert torch.equal(res.indices[ok], ref.indices[ok])
    assert res.stats_dict()["fallback_rows"] >= 1


@pytest.mark.parametrize("kind", ["uniform", "gauss"])
def test_one_dimensional_grid_with_a_huge_cell_count_stays_exact(kind):
    """d = 1 lets the grid grow to 2^20 cells, where the float32 cell assignment is off by up to ~0.2
    cells: the stop test's margin must scale with the grid.  Checked against the C oracle on
    sampled rows and against the SIMT pruner on a query slice."""
    from oracle import knn_oracle
    rng = np.random.default_rng(31)
    n = 300000
    x = (rng.uniform(-1000, 1000, (n, 1)) if kind == "uniform" else rng.standard_normal((n, 1)) * 50).astype(np.float32)
    xt = torch.from_numpy(x).cuda()
    res = mepol_b200.knn_device(xt, 5, algo=_lib.KNN_GRID)
    rows = np.sort(rng.choice(n, 2000, replace=False))
    ref_d, ref_i = knn_oracle.knn_c(x, 5, queries=x[rows], mode="kd_tree")
    got_i = res.indices[torch.from_numpy(rows).cuda()].cpu().numpy()
    got_d = res.distances[torch.from_numpy(rows).cuda()].cpu().numpy()
    assert_knn_matches(got_d, got_i, ref_d, ref_i, x, queries=x[rows], rtol=1e-12)
    q = xt[100000:116384]
    a = mepol_b200.knn_device(xt, 5, queries=q, algo=_lib.KNN_SIMT)
    b = mepol_b200.knn_device(xt, 5, queries=q, algo=_lib.KNN_GRID)
    assert torch.equal(a.distances, b.distances)
    assert torch.equal(a.indices, b.indices)
