"""Tuning aid: tensor-core vs SIMT pruner at small feature dimensions (device-only search time).
fp16 operands make the band ~8x tighter than bf16 did, so the AUTO threshold on d can come down."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mepol_b200
from mepol_b200 import _lib


def timed(x, k1, algo, reps=5):
    for _ in range(2):
        r = mepol_b200.knn_device(x, k1, algo=algo)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r = mepol_b200.knn_device(x, k1, algo=algo)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, r


if __name__ == "__main__":
    for dist in ("gauss", "uniform", "traj"):
        for n in (10000, 50000):
            for d in (2, 3, 4, 5, 6, 7, 8, 12, 16):
                rng = np.random.default_rng(d * 1000 + n)
                if dist == "gauss":
                    x = rng.standard_normal((n, d))
                elif dist == "uniform":
                    x = rng.uniform(-1, 1, (n, d))
                else:   # random-walk trajectories of 500 steps (neighbouring rows are close)
                    x = np.cumsum(rng.standard_normal((n // 500, 500, d)) * 0.05, axis=1).reshape(n, d)
                xt = torch.from_numpy(x.astype(np.float32)).cuda()
                ts, rs = timed(xt, 5, _lib.KNN_SIMT)
                tt, rt = timed(xt, 5, _lib.KNN_TC)
                same = torch.equal(rs.indices, rt.indices) and torch.equal(rs.distances, rt.distances)
                st = rt.stats_dict()
                print(f"{dist:8s} n={n:6d} d={d:2d}  simt {ts:8.3f} ms  tc {tt:8.3f} ms  same={same}  cand/row {st['candidates']/n:6.2f} "
                      f"ovf {st['tc_overflow_rows']:6d} fb {st['fallback_rows']:6d}", flush=True)
