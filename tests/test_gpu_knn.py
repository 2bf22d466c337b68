"""GPU parity of the exact k-NN search (through the C ABI) against the oracle and the sklearn
golden vectors.  Bar: indices bit-exact (exact ties as sets), distances within 1e-5 relative."""
import numpy as np
import pytest
import torch

from conftest import assert_knn_matches, golden_files, load_golden

pytestmark = pytest.mark.gpu

import mepol_b200
from mepol_b200 import _lib


def _algos_for(d, k1):
    out = [_lib.KNN_SIMT]
    lib = _lib.load()
    need = _lib._sz(0)
    if lib.mepol_knn_workspace_bytes(4096, 4096, d, k1, _lib.KNN_TC, __import__("ctypes").byref(need)) == 0:
        out.append(_lib.KNN_TC)
    return out


@pytest.mark.parametrize("path", golden_files("knn_"), ids=lambda p: p.split("knn_")[-1][:-4])
def test_matches_sklearn_golden(path):
    g = load_golden(path)
    x, k1 = g["x"], int(g["k1"])
    for algo in _algos_for(x.shape[1], k1):
        res = mepol_b200.knn_device(torch.from_numpy(x).cuda(), k1, algo=algo)
        assert_knn_matches(res.distances.cpu().numpy(), res.indices.cpu().numpy().astype(np.int64),
                           g["dist"], g["idx"].astype(np.int64), x=x)
        np.testing.assert_array_equal(res.rk.cpu().numpy(), res.distances[:, -1].cpu().numpy())


@pytest.mark.parametrize("n,d,k1,kind", [
    (20000, 29, 5, "normal"), (8192, 7, 5, "normal"), (12000, 2, 51, "uniform"), (5000, 24, 5, "normal"),
    (3000, 47, 5, "normal"), (257, 29, 5, "normal"), (6, 29, 5, "normal"), (5, 3, 5, "normal"),
    (4099, 29, 2, "normal"), (16384, 29, 5, "offset"), (10000, 29, 11, "traj")])
def test_matches_c_oracle_on_seeded_batches(n, d, k1, kind):
    from oracle import knn_oracle
    rng = np.random.default_rng(n + d)
    if kind == "normal":
        x = rng.standard_normal((n, d))
    elif kind == "uniform":
        x = rng.uniform(-6, 6, (n, d))
    elif kind == "offset":
        x = 0.2 * rng.standard_normal((n, d)) + 100.0 + 3 * np.arange(d)
    else:
        L = 500
        x = np.cumsum(0.05 * rng.standard_normal((n // L, L, d)), axis=1).reshape(-1, d)
    x = np.ascontiguousarray(x, dtype=np.float32)
    ref_d, ref_i = knn_oracle.knn_c(x, k1, mode="kd_tree")   # exact float64 distances
    for algo in _algos_for(d, k1):
        res = mepol_b200.knn_device(torch.from_numpy(x).cuda(), k1, algo=algo)
        assert_knn_matches(res.distances.cpu().numpy(), res.indices.cpu().numpy().astype(np.int64),
                           ref_d, ref_i, x=x, rtol=1e-12)


def test_duplicates_overflow_goes_through_exact_fallback():
    """'knn aliasing' (SURVEY Q2): hundreds of identical particles overflow the candidate band; the
    exact fallback must return the lowest-index duplicates at distance 0."""
    rng = np.random.default_rng(5)
    x = rng.standard_normal((3000, 2)).astype(np.float32)
    x[500:900] = x[500]                       # 400 exact copies (MountainCar wall state)
    from oracle import knn_oracle
    ref_d, ref_i = knn_oracle.knn_c(x, 5, mode="kd_tree")
    res = mepol_b200.knn_device(torch.from_numpy(x).cuda(), 5, algo=_lib.KNN_SIMT)
    st = res.stats_dict()
    assert st["fallback_rows"] >= 400
    idx = res.indices.cpu().numpy().astype(np.int64)
    assert_knn_matches(res.distances.cpu().numpy(), idx, ref_d, ref_i, x=x, rtol=1e-12)
    np.testing.assert_array_equal(idx[500:900], np.tile(np.arange(500, 505), (400, 1)))
    assert np.all(res.distances.cpu().numpy()[500:900] == 0.0)


def test_query_slice_matches_full_search_bitwise():
    """Row partitioning (multi-GPU sharding) must not change any row's result."""
    rng = np.random.default_rng(11)
    x = torch.from_numpy(rng.standard_normal((9000, 29)).astype(np.float32)).cuda()
    full = mepol_b200.knn_device(x, 5)
    for lo, hi in ((0, 4500), (4500, 9000), (1234, 1301)):
        part = mepol_b200.knn_device(x, 5, queries=x[lo:hi])
        assert torch.equal(part.indices, full.indices[lo:hi])
        assert torch.equal(part.distances, full.distances[lo:hi])
    other = x[100:200].clone()                 # non-aliasing query buffer
    part = mepol_b200.knn_device(x, 5, queries=other)
    assert torch.equal(part.indices, full.indices[100:200])


@pytest.mark.parametrize("d", [29, 7, 2])          # tensor-core, tensor-core (one K step), grid
def test_host_entry_points_agree(d):
    rng = np.random.default_rng(3)
    x = rng.standard_normal((5000, d)).astype(np.float32)
    d1, i1 = mepol_b200.knn(x, 4)
    d2, i2 = mepol_b200.knn_host_c(x, 4)
    assert i1.dtype == np.int64 and d1.dtype == np.float64 and i1.shape == (5000, 5)
    np.testing.assert_array_equal(i1, i2)
    np.testing.assert_array_equal(d1, d2)
    np.testing.assert_array_equal(i1[:, 0], np.arange(5000))
    with pytest.raises(ValueError):
        mepol_b200.knn(x[:3], 4)
    bad = x.copy()
    bad[7, d - 1] = np.nan
    with pytest.raises(ValueError):
        mepol_b200.knn(bad, 4)


def test_permutation_equivariance():
    rng = np.random.default_rng(21)
    x = rng.standard_normal((6000, 29)).astype(np.float32)
    perm = rng.permutation(6000)
    a = mepol_b200.knn_device(torch.from_numpy(x).cuda(), 5)
    b = mepol_b200.knn_device(torch.from_numpy(x[perm]).cuda(), 5)
    ia = a.indices.cpu().numpy()
    ib = b.indices.cpu().numpy()
    np.testing.assert_array_equal(perm[ib], ia[perm])
    np.testing.assert_array_equal(b.distances.cpu().numpy(), a.distances.cpu().numpy()[perm])


@pytest.mark.parametrize("n", [262144, 1048576])
def test_large_batch_properties(n):
    """Size-independent properties at a size the CPU oracle cannot brute-force in seconds: self is
    column 0 at distance 0, rows ascending, and a random sample of rows equals the oracle's answer
    for those query rows against the full set."""
    from oracle import knn_oracle
    rng = np.random.default_rng(99)
    x = rng.standard_normal((n, 29)).astype(np.float32)
    res = mepol_b200.knn_device(torch.from_numpy(x).cuda(), 5)
    idx = res.indices.cpu().numpy().astype(np.int64)
    dist = res.distances.cpu().numpy()
    np.testing.assert_array_equal(idx[:, 0], np.arange(n))
    assert np.all(dist[:, 0] == 0) and np.all(np.diff(dist, axis=1) >= 0)
    rows = rng.choice(n, 256 if n <= 262144 else 96, replace=False)
    ref_d, ref_i = knn_oracle.knn_c(x, 5, queries=x[rows], mode="kd_tree")
    np.testing.assert_array_equal(idx[rows], ref_i)
    np.testing.assert_allclose(dist[rows], ref_d, rtol=1e-12)
    # a handful of rows may overflow the bf16 band at this size; they go through the exact fallback
    assert res.stats_dict()["fallback_rows"] <= 64
