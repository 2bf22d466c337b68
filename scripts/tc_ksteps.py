"""Tensor-core search time for small d (one K=16 MMA step when d + 3 <= 16)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mepol_b200
from mepol_b200 import _lib
from tc_lowdim import timed

for n in (10000, 50000, 262144):
    for d in (7, 13, 14, 29):
        x = torch.from_numpy(np.random.default_rng(d).standard_normal((n, d)).astype(np.float32)).cuda()
        t, r = timed(x, 5, _lib.KNN_TC, reps=10 if n < 100000 else 3)
        s = mepol_b200.knn_device(x, 5, algo=_lib.KNN_SIMT) if n <= 50000 else None
        same = None if s is None else bool(torch.equal(s.indices, r.indices) and torch.equal(s.distances, r.distances))
        print(f"n={n:7d} d={d:2d}  tc {t:8.3f} ms  same_as_simt={same}", flush=True)
