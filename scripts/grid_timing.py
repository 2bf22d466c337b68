"""Tuning aid: grid vs SIMT search at the reference's 2-D configuration shapes (device-only time)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mepol_b200
from mepol_b200 import _lib
from tc_lowdim import timed

for name, n, d, k1, gen in (
        ("grid_world", 24000, 2, 51, "uniform"), ("grid_world 5x", 120000, 2, 51, "uniform"),
        ("mountain_car", 8000, 2, 5, "traj"), ("mountain_car 5x", 40000, 2, 5, "traj"),
        ("gauss 3d", 50000, 3, 5, "gauss"), ("gauss 2d 1M", 1000000, 2, 5, "gauss")):
    rng = np.random.default_rng(1)
    if gen == "uniform":
        x = rng.uniform(-6, 6, (n, d))
    elif gen == "gauss":
        x = rng.standard_normal((n, d))
    else:
        x = np.cumsum(rng.standard_normal((n // 400, 400, d)) * 0.01, axis=1).reshape(n, d)
    xt = torch.from_numpy(x.astype(np.float32)).cuda()
    tg, rg = timed(xt, k1, _lib.KNN_GRID, reps=10)
    if n <= 200000:
        ts, rs = timed(xt, k1, _lib.KNN_SIMT, reps=3)
        same = torch.equal(rs.indices, rg.indices) and torch.equal(rs.distances, rg.distances)
    else:
        ts, same = float("nan"), None
    print(f"{name:16s} n={n:7d} d={d} k1={k1:2d}  simt {ts:8.3f} ms  grid {tg:8.3f} ms  same={same}  "
          f"evaluated/row {rg.stats_dict()['candidates'] / n:7.1f} fb {rg.stats_dict()['fallback_rows']}", flush=True)
