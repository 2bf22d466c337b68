"""Tuning aid: tensor-core vs SIMT pruner at small batch sizes (device-only search time)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mepol_b200
from mepol_b200 import _lib
from tc_lowdim import timed

for d in (4, 7, 24, 29):
    for n in (512, 1024, 2048, 3072, 4096, 8192):
        x = torch.from_numpy(np.random.default_rng(n + d).standard_normal((n, d)).astype(np.float32)).cuda()
        ts, rs = timed(x, 5, _lib.KNN_SIMT, reps=20)
        tt, rt = timed(x, 5, _lib.KNN_TC, reps=20)
        same = torch.equal(rs.indices, rt.indices) and torch.equal(rs.distances, rt.distances)
        print(f"d={d:2d} n={n:5d}  simt {ts:7.3f} ms  tc {tt:7.3f} ms  same={same}  {rt.stats_dict()}", flush=True)
