"""Seeded random sweep of shapes and data kinds against the CPU oracle (GPU, through the C ABI):
any d in [1, 64], any k+1 in [1, 60], N from k+1 up to a few thousand, including integer-lattice data
(masses of exact ties), tight clusters far from the origin and exact duplicates."""
import math

import numpy as np
import pytest
import torch

from conftest import assert_knn_matches

pytestmark = pytest.mark.gpu

import mepol_b200
from mepol_b200 import _lib
from mepol_b200 import entropy as E
from oracle import iw_oracle as O
from oracle import knn_oracle


def _data(rng, n, d, kind):
    if kind == "normal":
        x = rng.standard_normal((n, d))
    elif kind == "lattice":
        x = rng.integers(-3, 4, (n, d)).astype(np.float64)          # exact ties everywhere
    elif kind == "clusters":
        c = rng.standard_normal((5, d)) * 50
        x = c[rng.integers(0, 5, n)] + 1e-3 * rng.standard_normal((n, d))
    elif kind == "dups":
        x = rng.standard_normal((n, d))
        m = max(1, n // 3)
        x[:m] = x[0]
    else:  # "scaled": wildly different column scales
        x = rng.standard_normal((n, d)) * (10.0 ** rng.integers(-3, 4, d))
    return np.ascontiguousarray(x, dtype=np.float32)


@pytest.mark.parametrize("seed", range(40))
def test_random_knn_case(seed):
    rng = np.random.default_rng(9000 + seed)
    d = int(rng.choice([1, 2, 3, 5, 7, 8, 13, 16, 24, 29, 31, 33, 47, 64]))
    k1 = int(rng.choice([1, 2, 5, 5, 5, 11, 12, 13, 25, 51, 60]))
    n = int(rng.integers(k1, 4000)) if seed % 4 else int(rng.integers(4096, 9000))
    kind = ["normal", "lattice", "clusters", "dups", "scaled"][seed % 5]
    x = _data(rng, n, d, kind)
    ref_d, ref_i = knn_oracle.knn_c(x, k1, mode="kd_tree")
    algos = [_lib.KNN_AUTO, _lib.KNN_SIMT]
    if d <= 29 and k1 <= 12:
        algos.append(_lib.KNN_TC)
    if d <= 3:
        algos.append(_lib.KNN_GRID)
    for algo in algos:
        res = mepol_b200.knn_device(torch.from_numpy(x).cuda(), k1, algo=algo)
        assert_knn_matches(res.distances.cpu().numpy(), res.indices.cpu().numpy().astype(np.int64), ref_d, ref_i,
                           x=x, rtol=1e-12, atol_self=1e-9)


@pytest.mark.parametrize("seed", range(24))
def test_random_knn_case_with_separate_queries(seed):
    """Queries that are not the fitted rows (what a multi-GPU shard or an evaluation batch passes),
    some of them far outside the fitted set's extent; every path must agree with the oracle and,
    bit for bit, with each other."""
    rng = np.random.default_rng(5000 + seed)
    d = int(rng.choice([1, 2, 3, 4, 6, 7, 16, 24, 29, 40]))
    k1 = int(rng.choice([1, 5, 5, 12, 30]))
    n = int(rng.integers(max(k1, 600), 7000))
    nq = int(rng.integers(1, 900))
    kind = ["normal", "clusters", "scaled", "lattice"][seed % 4]
    x = _data(rng, n, d, kind)
    q = _data(rng, nq, d, kind)
    if seed % 3 == 0:
        q[: max(1, nq // 10)] *= 40.0                        # far outside the box / fp16 range of the scaled operands
    ref_d, ref_i = knn_oracle.knn_c(x, k1, queries=q, mode="kd_tree")
    algos = [_lib.KNN_AUTO, _lib.KNN_SIMT]
    if d <= 29 and k1 <= 12:
        algos.append(_lib.KNN_TC)
    if d <= 3:
        algos.append(_lib.KNN_GRID)
    xt, qt = torch.from_numpy(x).cuda(), torch.from_numpy(q).cuda()
    first = None
    for algo in algos:
        res = mepol_b200.knn_device(xt, k1, queries=qt, algo=algo)
        assert_knn_matches(res.distances.cpu().numpy(), res.indices.cpu().numpy().astype(np.int64), ref_d, ref_i,
                           x=x, queries=q, rtol=1e-12, atol_self=1e-9)
        if first is None:
            first = res
        else:
            assert torch.equal(first.indices, res.indices) and torch.equal(first.distances, res.distances)


@pytest.mark.parametrize("seed", range(20))
def test_random_objective_case(seed):
    rng = np.random.default_rng(7000 + seed)
    T = int(rng.choice([1, 2, 7, 40, 600]))
    lens = rng.integers(1, 80 if T > 40 else 400, T)
    n = int(lens.sum())
    k = int(rng.choice([1, 4, 4, 9, 50]))
    k = min(k, max(1, n - 1))
    ns = int(rng.choice([2, 7, 24, 29]))
    eps = float(rng.choice([0.0, 0.0, 1e-15]))
    idx = rng.integers(0, n, (n, k + 1)).astype(np.int32)
    idx[:, 0] = np.arange(n)
    rk = rng.uniform(0.05, 3.0, n)
    lt, lb = rng.normal(0, 0.05, n), rng.normal(0, 0.05, n)
    G = math.gamma(ns / 2 + 1)
    B = 0.1234
    dist = np.zeros((n, k + 1)); dist[:, k] = rk
    H_ref = O.entropy(lt, lb, lens, dist, idx, k, G, B, ns, eps)
    kl_ref, _ = O.kl(lt, lb, lens, idx, k, eps)
    g_ref = O.entropy_grad_logp(lt, lb, lens, dist, idx, k, G, ns, eps)
    dev = torch.device("cuda")
    h = E.SearchHandle(torch.from_numpy(idx).to(dev), torch.from_numpy(rk).to(dev), k + 1)
    off, _ = mepol_b200.traj_offsets(lens, dev)
    ltt = torch.tensor(lt, dtype=torch.float64, device=dev, requires_grad=True)
    H, kl_raw, _ = mepol_b200.iw_entropy(ltt, torch.tensor(lb, dtype=torch.float64, device=dev), h, off, T, k, ns, G, B, eps)
    H.backward()
    np.testing.assert_allclose(H.item(), H_ref, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(max(kl_raw.item(), 0.0), kl_ref, rtol=1e-8, atol=1e-13)
    sc = np.abs(g_ref).max() + 1e-300
    np.testing.assert_allclose(ltt.grad.cpu().numpy() / sc, g_ref / sc, rtol=0, atol=1e-10)
