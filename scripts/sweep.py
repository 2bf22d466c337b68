"""BASELINE.json configs[4]: synthetic k-NN entropy scaling sweep N=64k..4M, d=29, k=4 on one B200.
Device-resident search (CUDA events) + IW fwd/bwd kernels; prints one JSON line per N."""
import json, math, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mepol_b200
from mepol_b200 import _lib

for n in [65536, 131072, 262144, 524288, 1048576, 2097152, 4194304]:
    rng = np.random.default_rng(0)
    x = torch.from_numpy(rng.standard_normal((n, 29), dtype=np.float32)).cuda()
    r = mepol_b200.knn_device(x, 5)
    torch.cuda.synchronize()
    reps = 3 if n <= 1048576 else 1
    _lib.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r = mepol_b200.knn_device(x, 5, workspace=r._workspace, out=(r.indices, r.distances, r.rk, r.stats))
    e1.record(); torch.cuda.synchronize()
    pr = float(np.mean(_lib.profile_read()))
    ms = e0.elapsed_time(e1) / reps
    ok = bool((r.indices[:, 0] == torch.arange(n, device="cuda", dtype=torch.int32)).all()) and bool((r.distances[:, 0] == 0).all())
    tf = 2.0 * n * n * 29 / (pr * 1e-3) / 1e12
    print(json.dumps({"N": n, "search_ms": round(ms, 3), "prune_ms": round(pr, 3), "states_per_s": round(n / (ms * 1e-3)),
                      "prune_TFLOPs_algorithmic": round(tf, 1), "frac_of_1367.7": round(tf / 1367.7, 3),
                      "self_col0_ok": ok, **r.stats_dict(), "workspace_MB": round(r._workspace.numel() / 2**20)}), flush=True)
    del r, x
    torch.cuda.empty_cache()
