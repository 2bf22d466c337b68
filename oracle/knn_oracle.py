"""CPU oracle for the k-NN search step -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates /root/reference/src/algorithms/mepol.py:190-192 (sklearn ``NearestNeighbors`` fit +
``kneighbors`` on the same matrix, ``n_neighbors=k+1``, self included).  Three entry points:

* :func:`knn_c`      -- the plain-C restatement in ``knn_oracle.c`` (OpenMP), any size.
* :func:`knn_numpy`  -- a float64 numpy restatement for small cases (independent of the C file).
* :func:`knn_sklearn`-- the reference's own three sklearn calls (the third-party library that owns
  the arithmetic; it is installed in the image, so this also runs on the GPU box).

Pinned against sklearn 1.9.0 by tests/test_oracle_knn.py using tests/golden/knn_*.npz.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libmepol_oracle.so")
_lib = None


def build(force=False):
    """Compile knn_oracle.c with gcc (see oracle/Makefile)."""
    src = os.path.join(_HERE, "knn_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "_build/libmepol_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(_SO)
        lib.mepol_oracle_knn.restype = ctypes.c_int
        lib.mepol_oracle_knn.argtypes = [
            ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int,
            ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        lib.mepol_oracle_num_threads.restype = ctypes.c_int
        _lib = lib
    return _lib


def sklearn_auto_mode(d):
    """'auto' resolution of sklearn 1.9.0 (sklearn/neighbors/_base.py:615-626): brute iff d > 15."""
    return "brute" if d > 15 else "kd_tree"


def knn_c(x, k1, queries=None, mode=None, n_threads=0):
    """(k1)-NN of every row of ``queries`` (default: x itself) among the rows of x.

    Returns (dist float64 [n_q,k1], idx int64 [n_q,k1]); mode 'brute' or 'kd_tree' (default: what
    sklearn's algorithm='auto' would pick for this d)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    q = x if queries is None else np.ascontiguousarray(queries, dtype=np.float32)
    if mode is None:
        mode = sklearn_auto_mode(x.shape[1])
    lib = _load()
    dist = np.empty((q.shape[0], k1), dtype=np.float64)
    idx = np.empty((q.shape[0], k1), dtype=np.int64)
    rc = lib.mepol_oracle_knn(q.ctypes.data, q.shape[0], x.ctypes.data, x.shape[0], x.shape[1],
                              k1, 0 if mode == "brute" else 1, dist.ctypes.data, idx.ctypes.data,
                              int(n_threads))
    if rc != 0:
        raise ValueError("Expected n_neighbors <= n_samples_fit (oracle rc=%d)" % rc)
    return dist, idx


def num_threads():
    return _load().mepol_oracle_num_threads()


def knn_numpy(x, k1, mode=None):
    """Small-case float64 numpy restatement (O(N^2) memory): ranks on exact float64 squared
    distances with ties broken by ascending index, i.e. the membership rule of _heap.pyx:46."""
    x32 = np.ascontiguousarray(x, dtype=np.float32)
    if mode is None:
        mode = sklearn_auto_mode(x32.shape[1])
    xd = x32.astype(np.float64)
    if mode == "brute":
        nrm = (xd * xd).sum(1)
        d2 = np.maximum(nrm[:, None] + (-2.0 * (xd @ xd.T)) + nrm[None, :], 0.0)
    else:
        d2 = ((xd[:, None, :] - xd[None, :, :]) ** 2).sum(-1)
    n = xd.shape[0]
    order = np.lexsort((np.broadcast_to(np.arange(n), (n, n)), d2), axis=1)[:, :k1]
    sel = np.take_along_axis(d2, order, axis=1)
    if mode == "brute":
        dist = np.sqrt(sel.astype(np.float32)).astype(np.float64)
    else:
        dist = np.sqrt(sel)
    return dist, order.astype(np.int64)


def knn_sklearn(x, k1, queries=None, n_jobs=None):
    """The reference's own call sequence (mepol.py:190-192) on the installed scikit-learn."""
    from sklearn.neighbors import NearestNeighbors
    nbrs = NearestNeighbors(n_neighbors=k1, metric="euclidean", algorithm="auto", n_jobs=n_jobs)
    nbrs.fit(x)
    return nbrs.kneighbors(x if queries is None else queries)
