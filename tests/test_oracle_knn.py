"""The k-NN oracle (oracle/knn_oracle.{c,py}) is pinned against sklearn outputs generated from the
reference's own call sequence (tests/golden/knn_*.npz, made by oracle/gen_golden.py)."""
import numpy as np
import pytest

from conftest import assert_knn_matches, golden_files, load_golden
from oracle import knn_oracle


@pytest.mark.parametrize("path", golden_files("knn_"), ids=lambda p: p.split("knn_")[-1][:-4])
def test_c_oracle_matches_sklearn_golden(path):
    g = load_golden(path)
    dist, idx = knn_oracle.knn_c(g["x"], int(g["k1"]), mode=str(g["method"]))
    assert_knn_matches(dist, idx, g["dist"], g["idx"], x=g["x"], rtol=1e-6)


@pytest.mark.parametrize("path", golden_files("knn_")[:6], ids=lambda p: p.split("knn_")[-1][:-4])
def test_numpy_oracle_matches_sklearn_golden(path):
    g = load_golden(path)
    dist, idx = knn_oracle.knn_numpy(g["x"], int(g["k1"]), mode=str(g["method"]))
    assert_knn_matches(dist, idx, g["dist"], g["idx"], x=g["x"], rtol=1e-6)


def test_auto_mode_matches_recorded_method():
    for path in golden_files("knn_"):
        g = load_golden(path)
        assert knn_oracle.sklearn_auto_mode(g["x"].shape[1]) == str(g["method"])


def test_live_sklearn_agrees_with_c_oracle():
    rng = np.random.default_rng(7)
    x = rng.standard_normal((900, 29)).astype(np.float32)
    d0, i0 = knn_oracle.knn_sklearn(x, 5)
    d1, i1 = knn_oracle.knn_c(x, 5)
    assert_knn_matches(d1, i1, d0, i0, x=x, rtol=1e-6)


def test_separate_queries_and_k_too_large():
    rng = np.random.default_rng(8)
    x = rng.standard_normal((300, 7)).astype(np.float32)
    q = x[40:90]
    d_all, i_all = knn_oracle.knn_c(x, 5)
    d_q, i_q = knn_oracle.knn_c(x, 5, queries=q)
    np.testing.assert_array_equal(i_q, i_all[40:90])
    np.testing.assert_array_equal(d_q, d_all[40:90])
    with pytest.raises(ValueError):
        knn_oracle.knn_c(x[:3], 5)
