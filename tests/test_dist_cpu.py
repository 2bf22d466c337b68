"""world_size-2 gloo run (CPU) of the multi-GPU search plumbing: row partition, all-gather of the
fitted set, all-gather of the per-shard results.  The CUDA search itself is replaced by the CPU
oracle through ShardedSearch's injection point, so this checks only the host/collective logic."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mepol_b200.dist import ShardedSearch, shard_bounds


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_search(k1):
    from oracle import knn_oracle

    def fn(full, queries):
        d, i = knn_oracle.knn_c(full.numpy(), k1, queries=queries.numpy(), mode="kd_tree", n_threads=1)
        return torch.from_numpy(i.astype(np.int32)), torch.from_numpy(d), torch.from_numpy(d[:, -1].copy())
    return fn


def _worker(rank, world, port, n, d, k1, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)
        x = torch.from_numpy(rng.standard_normal((n, d)).astype(np.float32))
        lo, hi = shard_bounds(n, world, rank)
        s = ShardedSearch(n, d, k1, torch.device("cpu"), search_fn=_oracle_search(k1))
        idx, rk, dst = s(x[lo:hi].contiguous(), gather_distances=True)
        torch.save({"idx": idx.clone(), "rk": rk.clone(), "dist": dst.clone()}, os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_shard_bounds():
    assert shard_bounds(1024, 4, 0) == (0, 256) and shard_bounds(1024, 4, 3) == (768, 1024)
    with pytest.raises(ValueError):
        shard_bounds(1001, 2, 0)


def test_two_rank_gloo_search_equals_single_rank(tmp_path):
    n, d, k1, world = 600, 7, 5, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, d, k1, str(tmp_path)), nprocs=world, join=True)
    from oracle import knn_oracle
    rng = np.random.default_rng(0)
    x = rng.standard_normal((n, d)).astype(np.float32)
    ref_d, ref_i = knn_oracle.knn_c(x, k1, mode="kd_tree")
    outs = [torch.load(os.path.join(tmp_path, f"r{r}.pt")) for r in range(world)]
    for o in outs:   # every rank holds the full, identical result
        np.testing.assert_array_equal(o["idx"].numpy(), ref_i)
        np.testing.assert_array_equal(o["dist"].numpy(), ref_d)
        np.testing.assert_array_equal(o["rk"].numpy(), ref_d[:, -1])


def test_single_rank_path_without_process_group():
    n, d, k1 = 300, 3, 5
    rng = np.random.default_rng(1)
    x = torch.from_numpy(rng.standard_normal((n, d)).astype(np.float32))
    s = ShardedSearch(n, d, k1, torch.device("cpu"), search_fn=_oracle_search(k1))
    idx, rk, dst = s(x, gather_distances=True)
    from oracle import knn_oracle
    ref_d, ref_i = knn_oracle.knn_c(x.numpy(), k1, mode="kd_tree")
    np.testing.assert_array_equal(idx.numpy(), ref_i)
    np.testing.assert_array_equal(dst.numpy(), ref_d)
