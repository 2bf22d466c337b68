"""Where the host-buffer search path (mepol_b200.search_particles) spends its time."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mepol_b200
from mepol_b200.knn import knn_device, to_device_f32, to_host


def T():
    torch.cuda.synchronize()
    return time.perf_counter()


for n, d, k in [(50000, 7, 4), (10000, 7, 4), (24000, 2, 50), (120000, 2, 50)]:
    x = np.random.default_rng(0).standard_normal((n, d)).astype(np.float32)
    mepol_b200.search_particles(x, k)
    for rep in range(3):
        t0 = T()
        ok = np.isfinite(x).all()
        ta = T()
        t1 = T()
        xd = to_device_f32(x, torch.device("cuda", 0))       # persistent pinned input block + async H2D
        t2 = T()
        res = knn_device(xd, k + 1)
        t3 = T()
        dist, idx = to_host(res)
        t4 = T()
        whole0 = T()
        mepol_b200.search_particles(x, k)
        whole1 = T()
        print(f"N={n} d={d} k={k}: isfinite {1e3*(ta-t0):.2f}  stage+h2d {1e3*(t2-t1):.2f}  search {1e3*(t3-t2):.2f}  "
              f"to_host {1e3*(t4-t3):.2f}  | search_particles {1e3*(whole1-whole0):.2f} ms", flush=True)
