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

    def _cuda_search(self, full, queries):
        res = knn_device(full, self.k1, queries=queries, algo=self.algo, workspace=self._ws, out=self._out)
        self._ws = res._workspace
        self._out = (res.indices, res.distances, res.rk, res.stats)
        return res.indices, res.distances, res.rk

    def __call__(self, shard, gather_distances=False):
        """shard: float32 [N/G, d] on this rank's device.  Returns (idx_full int32 [N,k1],
        rk_full float64 [N], dist_full or None) -- identical on every rank."""
        if self.world == 1:
            idx, dst, rk = self.search_fn(shard, shard)
            return idx, rk, (dst if gather_distances else None)
        dist.all_gather_into_tensor(self.full, shard.contiguous(), group=self.group)
        idx, dst, rk = self.search_fn(self.full, self.full[self.lo:self.hi])
        dist.all_gather_into_tensor(self.idx_full, idx.contiguous(), group=self.group)
        dist.all_gather_into_tensor(self.rk_full, rk.contiguous(), group=self.group)
        if gather_distances:
            if self.dist_full is None:
                self.dist_full = torch.empty((self.n, self.k1), dtype=torch.float64, device=self.device)
            dist.all_gather_into_tensor(self.dist_full, dst.contiguous(), group=self.group)
        return self.idx_full, self.rk_full, (self.dist_full if gather_distances else None)
