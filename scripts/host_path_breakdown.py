import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mepol_b200
from mepol_b200 import _lib
from mepol_b200.knn import knn_device
def T(): torch.cuda.synchronize(); return time.perf_counter()
for n, d, k in [(50000, 7, 4), (10000, 7, 4), (120000, 2, 50)]:
    x = np.random.default_rng(0).standard_normal((n, d)).astype(np.float32)
    mepol_b200.search_particles(x, k)
    for rep in range(2):
        t0 = T(); xt = torch.from_numpy(x).pin_memory(); t1 = T()
        xd = xt.to("cuda", non_blocking=True); t2 = T()
        res = knn_device(xd, k + 1); t3 = T()
        dist = res.distances.cpu(); idx = res.indices.to(torch.int64).cpu(); t4 = T()
        print(f"N={n} d={d} k={k}: pin {1e3*(t1-t0):.2f}  h2d {1e3*(t2-t1):.2f}  search {1e3*(t3-t2):.2f}  d2h {1e3*(t4-t3):.2f} ms  stats {res.stats_dict()}", flush=True)
