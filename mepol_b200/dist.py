"""Multi-GPU form of the search step: one process per GPU, query rows partitioned, the fitted
particle set all-gathered over NCCL/NVLink (north_star).

The reference has no distributed code at all (joblib processes for rollouts only,
/root/reference/src/algorithms/mepol.py:183-186); rollout workers map naturally to ranks: every rank
holds the particles its own workers collected (a contiguous row shard of next_states).

    rank g owns rows [g*N/G, (g+1)*N/G)                     (equal shards; N % G == 0)
    all_gather(shard)            -> full reference set on every rank       (the only exchange step)
    knn_device(full, queries=full[own rows])                -> idx/dist/rk for the own rows
    all_gather(idx, rk)          -> the owner rank (all ranks, it is 24 B/row) holds the full result

Row results do not depend on the partition (tests/test_gpu_knn.py::test_query_slice...), so the
output is bit-identical for every G.  The IW objective runs on the rank that owns the batch
(rank 0): it needs one global normaliser and a global gather, and costs microseconds.
Works with the `gloo` backend on CPU tensors for the partition/collective logic (tests), where the
search itself is injected.
"""
import mmap
import os
import uuid

import numpy as np
import torch
import torch.distributed as dist

from .knn import knn_device


def shard_bounds(n, world_size, rank):
    if n % world_size != 0:
        raise ValueError(f"N={n} must be divisible by the number of ranks ({world_size})")
    per = n // world_size
    return rank * per, (rank + 1) * per


class ShardedSearch:
    """Pre-allocated buffers for repeated sharded searches of a fixed shape."""

    def __init__(self, n, d, k1, device, group=None, search_fn=None, algo=0):
        self.n, self.d, self.k1, self.device, self.group, self.algo = n, d, k1, device, group, algo
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.lo, self.hi = shard_bounds(n, self.world, self.rank)
        self.search_fn = search_fn or self._cuda_search
        self.full = torch.empty((n, d), dtype=torch.float32, device=device)
        self.idx_full = torch.empty((n, k1), dtype=torch.int32, device=device)
        self.rk_full = torch.empty((n,), dtype=torch.float64, device=device)
        self.dist_full = None
        self._ws = None
        self._out = None
        self.last_idx = self.last_rk = self.last_dist = None
        self.own_idx = self.own_dist = None

    def _cuda_search(self, full, queries):
        res = knn_device(full, self.k1, queries=queries, algo=self.algo, workspace=self._ws, out=self._out)
        self._ws = res._workspace
        self._out = (res.indices, res.distances, res.rk, res.stats)
        return res.indices, res.distances, res.rk

    def __call__(self, shard, gather_distances=False, _marks=None):
        """shard: float32 [N/G, d] on this rank's device.  Returns (idx_full int32 [N,k1],
        rk_full float64 [N], dist_full or None) -- identical on every rank."""
        mark = _marks.append if _marks is not None else (lambda _e: None)

        def ev():
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            return e
        if _marks is not None:
            mark(ev())
        if self.world == 1:
            if _marks is not None:
                mark(ev())
            idx, dst, rk = self.search_fn(shard, shard)
            self.own_idx, self.own_dist = idx, dst
            if _marks is not None:
                mark(ev())
                mark(ev())
            self.full = shard
            self.last_idx, self.last_rk, self.last_dist = idx, rk, dst
            return idx, rk, (dst if gather_distances else None)
        dist.all_gather_into_tensor(self.full, shard.contiguous(), group=self.group)
        if _marks is not None:
            mark(ev())
        idx, dst, rk = self.search_fn(self.full, self.full[self.lo:self.hi])
        self.own_idx, self.own_dist = idx, dst           # this rank's rows (for SharedHostResults.push)
        if _marks is not None:
            mark(ev())
        dist.all_gather_into_tensor(self.idx_full, idx.contiguous(), group=self.group)
        dist.all_gather_into_tensor(self.rk_full, rk.contiguous(), group=self.group)
        if gather_distances:
            if self.dist_full is None:
                self.dist_full = torch.empty((self.n, self.k1), dtype=torch.float64, device=self.device)
            dist.all_gather_into_tensor(self.dist_full, dst.contiguous(), group=self.group)
        if _marks is not None:
            mark(ev())
        self.last_idx, self.last_rk = self.idx_full, self.rk_full
        self.last_dist = self.dist_full if gather_distances else None
        return self.idx_full, self.rk_full, (self.dist_full if gather_distances else None)

    def timed_phases(self, shard, reps=3):
        """Per-phase milliseconds of one sharded search on this rank (CUDA events on the launching stream):
        all-gather of the particles, the search of the own rows (with the pruning kernel's share from the
        library's own event pair), all-gather of the results."""
        from . import _lib
        out = {"gather_particles_ms": 0.0, "search_own_rows_ms": 0.0, "prune_kernel_ms": 0.0, "gather_results_ms": 0.0}
        self(shard)
        for _ in range(reps):
            marks = []
            _lib.profile_enable(True)
            self(shard, _marks=marks)
            torch.cuda.synchronize(self.device)
            pr = _lib.profile_read(4)
            _lib.profile_enable(False)
            out["gather_particles_ms"] += marks[0].elapsed_time(marks[1]) / reps
            out["search_own_rows_ms"] += marks[1].elapsed_time(marks[2]) / reps
            out["gather_results_ms"] += marks[2].elapsed_time(marks[3]) / reps
            out["prune_kernel_ms"] += (pr[0] if pr else 0.0) / reps
        return out


class SharedHostResults:
    """Host-side landing zone of a sharded search on ONE box: sklearn's output arrays (indices int64
    [N, k1], distances float64 [N, k1], mepol.py:192) live in one POSIX shared-memory block that every
    rank maps and page-locks (cudaHostRegister), and each GPU copies the rows it searched straight into
    it -- G concurrent PCIe transfers of N/G rows each instead of gathering everything on the owner's
    GPU and pushing 16 * k1 * N bytes through one link.  The owner rank (0) reads the arrays after the
    barrier.  With one rank this degenerates to a pinned block.  No device-side collective is involved
    for the distances at all (the objective only needs idx + r_k, which ShardedSearch gathers)."""

    def __init__(self, n, k1, device=None, group=None):
        self.n, self.k1, self.group = n, k1, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.device = device
        self.lo, self.hi = shard_bounds(n, self.world, self.rank)
        nbytes = 16 * n * k1
        self._registered = False
        self._map = self._file = None
        if self.world == 1:
            # one rank: a plain page-locked block does it (no shared mapping needed)
            pin = device is not None and torch.device(device).type == "cuda"
            buf = torch.empty(nbytes, dtype=torch.uint8, pin_memory=pin)
        else:
            name = [None]
            if self.rank == 0:
                st = os.statvfs("/dev/shm")
                if st.f_bavail * st.f_frsize < nbytes + (8 << 20):
                    name[0] = ""                 # every rank must learn about the failure (no hang)
                else:
                    name[0] = f"/dev/shm/mepol_b200_{os.getpid()}_{uuid.uuid4().hex[:12]}"
                    with open(name[0], "wb") as f:
                        f.truncate(nbytes)
            dist.broadcast_object_list(name, src=0, group=group)
            if not name[0]:
                raise RuntimeError(f"/dev/shm has less than {nbytes >> 20} MiB free for the shared result block")
            self._path = name[0]
            self._file = open(self._path, "r+b")
            self._map = mmap.mmap(self._file.fileno(), nbytes)
            dist.barrier(group=group)
            if self.rank == 0:
                os.unlink(self._path)               # the mapping keeps the memory alive; nothing is left behind
            buf = torch.frombuffer(self._map, dtype=torch.uint8)
            if device is not None and torch.device(device).type == "cuda":
                rc = torch.cuda.cudart().cudaHostRegister(buf.data_ptr(), nbytes, 0)
                if int(rc) != 0:
                    raise RuntimeError(f"cudaHostRegister of the shared result block failed (rc={int(rc)})")
                self._registered = True
        self._buf = buf
        self.indices = buf[:8 * n * k1].view(torch.int64).view(n, k1)
        self.distances = buf[8 * n * k1:].view(torch.float64).view(n, k1)

    def push(self, idx_rows, dist_rows):
        """This rank's rows (device or host tensors, [N/G, k1]) -> the shared arrays, asynchronously on
        the current stream; the int32 -> int64 widening happens on the device in front of the copy."""
        self.indices[self.lo:self.hi].copy_(idx_rows, non_blocking=True)
        self.distances[self.lo:self.hi].copy_(dist_rows, non_blocking=True)

    def wait(self):
        """All ranks' copies have landed (stream sync + barrier)."""
        if self.device is not None and torch.device(self.device).type == "cuda":
            torch.cuda.current_stream(self.device).synchronize()
        if self.world > 1:
            dist.barrier(group=self.group)

    def close(self):
        if self._buf is None:
            return
        if self._registered:
            torch.cuda.cudart().cudaHostUnregister(self._buf.data_ptr())
        self.indices = self.distances = self._buf = None
        if self._map is not None:
            try:
                self._map.close()
            except BufferError:
                pass
            self._file.close()
