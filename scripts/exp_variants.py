"""A/B of pruning-kernel variants (MEPOL_TC_VARIANT is read at every launch): prune-kernel time from the
library's own CUDA events, results compared bitwise with variant 0.
    python scripts/exp_variants.py 0,1 [N ...]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mepol_b200
from mepol_b200 import _lib

variants = sys.argv[1].split(",") if len(sys.argv) > 1 else ["0", "1"]
sizes = [int(a) for a in sys.argv[2:]] or [131072, 1048576]
d, k1 = int(os.environ.get("EXP_D", "29")), 5
for n in sizes:
    rng = np.random.default_rng(0)
    x = torch.from_numpy(rng.standard_normal((n, d), dtype=np.float32)).cuda()
    base = None
    extra_set = []
    for v in variants:
        # "3;MEPOL_TC_WARM_TILES=64;X=Y" = variant 3 with extra environment settings for this run
        parts = v.split(";")
        os.environ["MEPOL_TC_VARIANT"] = parts[0]
        for kv in extra_set:
            os.environ.pop(kv, None)
        extra_set = []
        for kv in parts[1:]:
            key, val = kv.split("=")
            os.environ[key] = val
            extra_set.append(key)
        r = mepol_b200.knn_device(x, k1, algo=_lib.KNN_TC)
        torch.cuda.synchronize()
        reps = 5 if n <= 1048576 else 2
        _lib.profile_enable(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            r = mepol_b200.knn_device(x, k1, algo=_lib.KNN_TC, workspace=r._workspace,
                                      out=(r.indices, r.distances, r.rk, r.stats))
        e1.record()
        torch.cuda.synchronize()
        pr = _lib.profile_read()
        _lib.profile_enable(False)
        if base is None:
            base = (r.indices.clone(), r.distances.clone())
            same = True
        else:
            same = torch.equal(base[0], r.indices) and torch.equal(base[1], r.distances)
        tf = 2.0 * n * n * d / (min(pr) * 1e-3) / 1e12
        print(json.dumps({"N": n, "variant": v, "prune_ms_min": round(min(pr), 3), "prune_ms_mean": round(float(np.mean(pr)), 3),
                          "search_ms": round(e0.elapsed_time(e1) / reps, 3), "TF": round(tf, 1),
                          "frac_sustained": round(tf / 1367.7, 3), "same_as_v0": same, **r.stats_dict()}), flush=True)
        del r
    del x, base
    torch.cuda.empty_cache()
