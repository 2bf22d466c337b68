"""BASELINE.json configs[4]: synthetic k-NN entropy scaling sweep N=64k..4M, d=29, k=4 at 1/2/4/8 B200.

    python scripts/sweep.py                         # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 scripts/sweep.py

One JSON line per N (rank 0): step time (sharded search + reverse CSR + IW forward/backward on rank 0, CUDA
events, max over ranks), the pruning kernel's share and roofline fraction, sampled-row parity against the
C oracle (exact float64 brute force of oracle/knn_oracle.c on 256 random query rows of the very array that
was searched), and for N <= 262144 on one GPU the reference's sklearn search timed on the host cores.
"""
import json
import math
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mepol_b200
from mepol_b200 import _lib
from mepol_b200 import entropy as E
from mepol_b200.dist import ShardedSearch, shard_bounds

SIZES = [65536, 131072, 262144, 524288, 1048576, 2097152, 4194304]
D, K, L = 29, 4, 500


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    sizes = [int(a) for a in sys.argv[1:]] or SIZES
    peak = 1367.7
    try:
        with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) as f:
            peak = float(json.load(f)["bf16_tflops_sustained"])
    except Exception:
        pass
    import scipy.special
    G = float(scipy.special.gamma(D / 2 + 1))
    B = float(math.log(K) - scipy.special.digamma(K))
    for n in sizes:
        rng = np.random.default_rng(0)
        x = rng.standard_normal((n, D), dtype=np.float32)
        lt = torch.from_numpy(rng.normal(0.0, 0.01, n).astype(np.float32)).to(dev).requires_grad_(True)
        lb = torch.zeros(n, dtype=torch.float32, device=dev)
        lens = [L] * (n // L) + ([n % L] if n % L else [])
        off, _ = mepol_b200.traj_offsets(lens, dev)
        lo, hi = shard_bounds(n, world, rank)
        shard = torch.from_numpy(x[lo:hi]).to(dev)
        searcher = ShardedSearch(n, D, K + 1, dev)

        def step():
            idx_full, rk_full, _ = searcher(shard)
            if rank == 0:
                h = E.SearchHandle(idx_full, rk_full, K + 1)
                h.csr()
                H, _, _ = mepol_b200.iw_entropy(lt, lb, h, off, len(lens), K, D, G, B, 0.0)
                lt.grad = None
                H.backward()

        for _ in range(2):
            step()
        reps = 5 if n <= 1048576 else 2
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        _lib.profile_enable(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            step()
        e1.record()
        torch.cuda.synchronize()
        pr = _lib.profile_read()
        _lib.profile_enable(False)
        ms = e0.elapsed_time(e1) / reps
        t = torch.tensor([ms, float(np.mean(pr))], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, prune = float(t[0]), float(t[1])
        if rank == 0:
            from oracle import knn_oracle
            rows = np.sort(np.random.default_rng(1).choice(n, 256, replace=False))
            ref_d, ref_i = knn_oracle.knn_c(x, K + 1, queries=x[rows], mode="kd_tree")
            got_i = searcher.last_idx[torch.from_numpy(rows).to(dev)].cpu().numpy().astype(np.int64)
            got_rk = searcher.last_rk[torch.from_numpy(rows).to(dev)].cpu().numpy()
            mism = int((got_i != ref_i).any(axis=1).sum())
            rk_err = float(np.max(np.abs(got_rk - ref_d[:, K]) / ref_d[:, K]))
            tf = 2.0 * (hi - lo) * n * D / (prune * 1e-3) / 1e12
            line = {"N": n, "n_gpus": world, "step_ms": round(ms, 3), "prune_ms": round(prune, 3),
                    "states_per_s": round(n / (ms * 1e-3)), "prune_TFLOPs_algorithmic_per_gpu": round(tf, 1),
                    "frac_of_sustained_peak": round(tf / peak, 3),
                    "oracle_rows_checked": 256, "oracle_row_mismatch": mism, "rk_max_rel_err": rk_err}
            if world == 1:
                line["search_stats"] = {k_: int(v) for k_, v in zip(("candidates", "fallback_rows", "compactions"),
                                                                   searcher._out[3].cpu().tolist()[:3])}
            if world == 1 and n <= 262144 and not os.environ.get("SWEEP_NO_CPU"):
                t0 = time.perf_counter()
                sk_d, sk_i = knn_oracle.knn_sklearn(x, K + 1, n_jobs=os.cpu_count())
                t_cpu = time.perf_counter() - t0
                full_i = searcher.last_idx.cpu().numpy().astype(np.int64)
                line["sklearn_full_search_s"] = round(t_cpu, 2)
                line["sklearn_cores"] = os.cpu_count()
                line["sklearn_states_per_s"] = round(n / t_cpu)
                line["sklearn_row_mismatch_all_rows"] = int((full_i != sk_i).any(axis=1).sum())
            print(json.dumps(line), flush=True)
        del searcher, shard, lt, lb
        torch.cuda.empty_cache()
        if world > 1:
            dist.barrier()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
