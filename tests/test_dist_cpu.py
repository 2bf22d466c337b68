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

from mepol_b200.dist import ShardedSearch, SharedHostResults, shard_bounds


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


def _sink_worker(rank, world, port, n, k1, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sink = SharedHostResults(n, k1, device=None)        # CPU: same mapping logic, no page-locking
        lo, hi = shard_bounds(n, world, rank)
        rows = torch.arange(lo, hi, dtype=torch.int32).reshape(-1, 1).repeat(1, k1) * 10 + torch.arange(k1, dtype=torch.int32)
        sink.push(rows, rows.to(torch.float64) / 7.0)
        sink.wait()
        if rank == 0:
            torch.save({"idx": sink.indices.clone(), "dist": sink.distances.clone()}, os.path.join(out_dir, "sink.pt"))
        dist.barrier()
        sink.close()
    finally:
        dist.destroy_process_group()


def test_shared_host_results_collects_every_ranks_rows(tmp_path):
    """Each rank writes the rows it searched into the owner's shared host arrays (the e2e path of
    bench.py at N > 1): the owner sees all of them, widened to sklearn's int64 / float64."""
    n, k1, world = 4000, 5, 2
    port = _free_port()
    mp.spawn(_sink_worker, args=(world, port, n, k1, str(tmp_path)), nprocs=world, join=True)
    got = torch.load(os.path.join(tmp_path, "sink.pt"))
    expect = torch.arange(n, dtype=torch.int64).reshape(-1, 1).repeat(1, k1) * 10 + torch.arange(k1)
    assert got["idx"].dtype == torch.int64 and torch.equal(got["idx"], expect)
    assert torch.equal(got["dist"], expect.to(torch.float64) / 7.0)
    assert not [f for f in os.listdir("/dev/shm") if f.startswith("mepol_b200_")], "shared block left behind"
