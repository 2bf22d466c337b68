"""Bring-up cycle counters at mid-size batches (build with MEPOL_NVCC_EXTRA=-DMEPOL_BRINGUP, run with MEPOL_TC_PROF=1)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mepol_b200
from mepol_b200 import _lib

for n in (65536, 262144):
    x = torch.from_numpy(np.random.default_rng(0).standard_normal((n, 29)).astype(np.float32)).cuda()
    for _ in range(2):
        r = mepol_b200.knn_device(x, 5, algo=_lib.KNN_TC)
    torch.cuda.synchronize()
    print("N", n, r.stats_dict(), flush=True)
