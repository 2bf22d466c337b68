import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def pytest_sessionstart(session):
    """Make the suite self-contained: build the C-ABI library (nvcc cross-compiles sm_100a without a
    GPU) and the C oracle if they are missing or stale.  Both builds are incremental no-ops otherwise."""
    import shutil
    from mepol_b200 import build as _build
    if shutil.which("nvcc") or os.path.exists(_build.NVCC):
        _build.build()
    elif not os.path.exists(_build.LIB):
        raise RuntimeError("libmepol_b200.so is missing and nvcc is not available to build it")
    from oracle import knn_oracle
    knn_oracle.build()


def golden_files(prefix):
    return sorted(glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def load_golden(path):
    with np.load(path, allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def exact_d2(x, rows, cols):
    xd = np.asarray(x, np.float64)
    return ((xd[rows] - xd[cols]) ** 2).sum(-1)


def assert_knn_matches(dist, idx, ref_dist, ref_idx, x, queries=None, rtol=1e-5, atol_self=1e-6):
    """Parity contract of SURVEY.md Q2/Q3:
      * every index equals the reference's, EXCEPT where the two candidates are at exactly the same
        float64 distance from the query (duplicates / "knn aliasing"): sklearn's order and boundary
        membership among exact ties is a heap/tree-traversal artefact, so either is accepted;
      * no index repeats inside a row;
      * distances agree within rtol relative (north_star: 1e-5); the near-zero self column is
        compared absolutely (sklearn's expanded float64 form returns ~1e-7*||x|| there, not 0)."""
    dist, idx, ref_dist, ref_idx = map(np.asarray, (dist, idx, ref_dist, ref_idx))
    assert dist.shape == ref_dist.shape and idx.shape == ref_idx.shape
    q = np.asarray(x if queries is None else queries, np.float64)
    xd = np.asarray(x, np.float64)
    bad_r, bad_c = np.nonzero(idx != ref_idx)
    if bad_r.size:
        mine = ((q[bad_r] - xd[idx[bad_r, bad_c]]) ** 2).sum(-1)
        theirs = ((q[bad_r] - xd[ref_idx[bad_r, bad_c]]) ** 2).sum(-1)
        np.testing.assert_array_equal(mine, theirs, err_msg="index mismatch that is not an exact tie")
    srt = np.sort(idx, axis=1)
    assert np.all(srt[:, 1:] != srt[:, :-1]), "repeated neighbour index inside a row"
    scale = float(np.sqrt(np.max(np.sum(xd ** 2, 1))))
    big = ref_dist > atol_self * max(scale, 1.0)
    np.testing.assert_allclose(dist[big], ref_dist[big], rtol=rtol)
    assert np.all(np.abs(dist[~big] - ref_dist[~big]) <= atol_self * max(scale, 1.0))
