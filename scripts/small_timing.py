import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mepol_b200
from mepol_b200 import _lib
for (n, d, k1) in [(8000, 2, 5), (24000, 2, 51), (24000, 2, 5), (10000, 7, 5), (10000, 24, 5), (50000, 7, 5), (120000, 2, 51)]:
    rng = np.random.default_rng(1)
    x = torch.from_numpy(rng.standard_normal((n, d)).astype(np.float32)).cuda()
    r = mepol_b200.knn_device(x, k1)
    torch.cuda.synchronize()
    _lib.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        r = mepol_b200.knn_device(x, k1, workspace=r._workspace, out=(r.indices, r.distances, r.rk, r.stats))
    e1.record(); torch.cuda.synchronize()
    pr = _lib.profile_read()
    print(f"N={n} d={d} k1={k1}: search {e0.elapsed_time(e1)/5:.3f} ms, prune {np.mean(pr):.3f} ms, stats {r.stats_dict()}", flush=True)
