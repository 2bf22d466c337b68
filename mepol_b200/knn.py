"""Host side of the exact k-NN search: the replacement for the three sklearn calls at
/root/reference/src/algorithms/mepol.py:190-192.

`knn(next_states, k)` has the meaning of
    NearestNeighbors(n_neighbors=k+1, metric='euclidean', algorithm='auto').fit(X).kneighbors(X)
and returns (distances float64 [N,k+1], indices int64 [N,k+1]) as numpy arrays, like sklearn.
`knn_device` is the same on tensors already resident in HBM (what the drop-in and bench use).
torch is used for device memory and streams only; every kernel is in libmepol_b200.so.
"""
import threading

import numpy as np
import torch

from . import _lib


def _stream_ptr(device):
    return torch.cuda.current_stream(device).cuda_stream


class KnnResult:
    """Device-resident search result.  `indices`/`distances` follow sklearn's layout; `rk` is the
    contiguous k-th-neighbour distance column the objective consumes (mepol.py:150)."""

    def __init__(self, indices, distances, rk, stats, k1, n_ref):
        self.indices = indices          # int32 [n_q, k1]
        self.distances = distances      # float64 [n_q, k1]
        self.rk = rk                    # float64 [n_q]
        self.stats = stats              # int64 [8] device
        self.k1 = k1
        self.n_ref = n_ref

    def stats_dict(self):
        s = self.stats.cpu().tolist()
        return {"candidates": s[0], "fallback_rows": s[1], "compactions": s[2], "tc_overflow_rows": s[3]}


def knn_device(ref, k1, queries=None, algo=_lib.KNN_AUTO, workspace=None, out=None):
    """Exact (k1)-NN of `queries` (default: ref itself, self included) among the rows of `ref`.

    ref: float32 CUDA tensor [n_ref, d], contiguous.  queries: float32 CUDA [n_q, d] (may be a row
    slice of ref).  Returns KnnResult.  Raises ValueError when k1 > n_ref (sklearn's behaviour)."""
    if not (isinstance(ref, torch.Tensor) and ref.is_cuda):
        raise _lib.MepolError("knn_device needs CUDA tensors (there is no CPU fallback)")
    if ref.dtype != torch.float32 or ref.dim() != 2 or not ref.is_contiguous():
        raise ValueError("ref must be a contiguous float32 [n, d] tensor")
    q = ref if queries is None else queries
    if q.dtype != torch.float32 or q.dim() != 2 or not q.is_contiguous() or q.shape[1] != ref.shape[1]:
        raise ValueError("queries must be a contiguous float32 [n_q, d] tensor with the same d")
    if q.device != ref.device:
        raise ValueError("queries and ref must live on the same device")
    n_ref, d = ref.shape
    n_q = q.shape[0]
    lib = _lib.load()
    with torch.cuda.device(ref.device):
        need = _lib.knn_workspace_bytes(n_q, n_ref, d, k1, algo)
        if workspace is None or workspace.numel() < need:
            workspace = torch.empty(need, dtype=torch.uint8, device=ref.device)
        if out is None:
            idx = torch.empty((n_q, k1), dtype=torch.int32, device=ref.device)
            dist = torch.empty((n_q, k1), dtype=torch.float64, device=ref.device)
            rk = torch.empty((n_q,), dtype=torch.float64, device=ref.device)
            stats = torch.zeros((_lib.KNN_NSTATS,), dtype=torch.int64, device=ref.device)
        else:
            idx, dist, rk, stats = out
        rc = lib.mepol_knn_search_f32(ref.data_ptr(), n_ref, q.data_ptr(), n_q, d, k1, algo,
                                      idx.data_ptr(), dist.data_ptr(), rk.data_ptr(),
                                      workspace.data_ptr(), workspace.numel(), stats.data_ptr(),
                                      _stream_ptr(ref.device))
        _lib.check(rc, "mepol_knn_search_f32")
    res = KnnResult(idx, dist, rk, stats, k1, n_ref)
    res._workspace = workspace  # keep alive until the stream has consumed it
    return res


class _PinnedStage:
    """One persistent pinned host block (grown geometrically) + the event of the last transfer that
    used it.  Allocating pinned memory per call costs milliseconds -- more than a whole search at the
    reference's batch sizes -- and a pageable source/destination makes the driver stage the copy itself."""

    def __init__(self):
        self.buf = None
        self.busy = None
        self.lock = threading.Lock()

    def take(self, nbytes):
        if self.busy is not None:          # the previous transfer through this block must have drained
            self.busy.synchronize()
            self.busy = None
        if self.buf is None or self.buf.numel() < nbytes:
            self.buf = torch.empty(max(nbytes, 1 << 20, 0 if self.buf is None else 2 * self.buf.numel()),
                                   dtype=torch.uint8, pin_memory=True)
        return self.buf

    def mark(self, device):
        self.busy = torch.cuda.Event()
        self.busy.record(torch.cuda.current_stream(device))


_STAGE_IN = _PinnedStage()


def to_device_f32(x, device):
    """Host float32 [n, d] numpy array -> device tensor through the persistent pinned input block
    (replaces the per-call `torch.from_numpy(x).pin_memory()` of round 1: rollout output staging,
    reference mepol.py:187,195-200)."""
    nbytes = x.nbytes
    with _STAGE_IN.lock:
        pin = _STAGE_IN.take(nbytes)[:nbytes].view(torch.float32).view(x.shape)
        pin.copy_(torch.from_numpy(x))                     # host memcpy into pinned memory
        xt = torch.empty(x.shape, dtype=torch.float32, device=device)
        xt.copy_(pin, non_blocking=True)
        _STAGE_IN.mark(device)
    return xt


def to_host(res):
    """(distances float64, indices int64) of a KnnResult as fresh CPU tensors owned by the caller.
    The results are copied device->host straight into caller-owned PINNED tensors (torch's caching host
    allocator recycles the blocks of earlier results once the caller dropped them, so in steady state no
    cudaHostAlloc happens), the int32 -> int64 widening of sklearn's index dtype runs on the device in
    front of the copy: no pageable staging and no host-side memcpy (at GridWorld's [24000, 51] outputs
    the host copies were 2/3 of the whole call)."""
    dev = res.indices.device
    d_out = torch.empty(res.distances.shape, dtype=torch.float64, pin_memory=True)
    i_out = torch.empty(res.indices.shape, dtype=torch.int64, pin_memory=True)
    d_out.copy_(res.distances, non_blocking=True)
    i_out.copy_(res.indices, non_blocking=True)
    torch.cuda.current_stream(dev).synchronize()
    return d_out, i_out


def knn(next_states, k, algo=_lib.KNN_AUTO, device=None):
    """Drop-in for mepol.py:190-192.  next_states: float32-convertible array [N, ns] on the HOST.
    Returns (distances float64 [N,k+1], indices int64 [N,k+1]) numpy arrays."""
    x = np.ascontiguousarray(next_states, dtype=np.float32)
    if x.ndim != 2:
        raise ValueError("Expected 2D array")
    if not np.isfinite(x).all():
        raise ValueError("Input contains NaN or infinity")   # sklearn check_array behaviour
    if not torch.cuda.is_available():
        raise _lib.MepolError("mepol_b200.knn needs a CUDA device (there is no CPU fallback)")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    with torch.cuda.device(dev):
        xt = to_device_f32(x, dev)
        res = knn_device(xt, k + 1, algo=algo)
        dist, idx = to_host(res)
    return dist.numpy(), idx.numpy()


def knn_host_c(next_states, k, algo=_lib.KNN_AUTO):
    """Same contract through the pure C host entry point mepol_knn_search_host_f32 (no torch)."""
    x = np.ascontiguousarray(next_states, dtype=np.float32)
    n, d = x.shape
    idx = np.empty((n, k + 1), dtype=np.int64)
    dist = np.empty((n, k + 1), dtype=np.float64)
    rc = _lib.load().mepol_knn_search_host_f32(x.ctypes.data, n, d, k + 1, algo, idx.ctypes.data, dist.ctypes.data)
    _lib.check(rc, "mepol_knn_search_host_f32")
    return dist, idx
