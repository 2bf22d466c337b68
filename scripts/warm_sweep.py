"""Tuning aid: search time at N=1M/d=29/k1=5 (and a few smaller sizes) against MEPOL_TC_WARM_TILES.
Run one value per process: MEPOL_TC_WARM_TILES=<w> python scripts/warm_sweep.py"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mepol_b200
from mepol_b200 import _lib

for n in (1048576, 131072, 32768):
    x = torch.from_numpy(np.random.default_rng(0).standard_normal((n, 29)).astype(np.float32)).cuda()
    for _ in range(2):
        r = mepol_b200.knn_device(x, 5, algo=_lib.KNN_TC)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3 if n > 500000 else 10
    e0.record()
    for _ in range(reps):
        r = mepol_b200.knn_device(x, 5, algo=_lib.KNN_TC)
    e1.record()
    torch.cuda.synchronize()
    print(os.environ.get("MEPOL_TC_WARM_TILES", "default"), n, "%.3f ms" % (e0.elapsed_time(e1) / reps), r.stats_dict(), flush=True)
