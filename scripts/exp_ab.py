"""Interleaved A/B of pruning-kernel settings (a power-capped board drifts by several per cent between back-to-back
runs of the SAME kernel, so variants alternate launch by launch and are compared on their means).
    python scripts/exp_ab.py "3|3;MEPOL_TC2_ISSUER=1" N [rounds]
Each variant = MEPOL_TC_VARIANT value followed by ;ENV=VALUE settings.  Results are compared bitwise with the first."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mepol_b200
from mepol_b200 import _lib

variants = sys.argv[1].split("|")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1048576
rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 8
d, k1 = int(os.environ.get("EXP_D", "29")), 5
x = torch.from_numpy(np.random.default_rng(0).standard_normal((n, d), dtype=np.float32)).cuda()
managed = set()


def setenv(v):
    parts = v.split(";")
    for key in managed:
        os.environ.pop(key, None)
    os.environ["MEPOL_TC_VARIANT"] = parts[0]
    for kv in parts[1:]:
        key, val = kv.split("=")
        os.environ[key] = val
        managed.add(key)


setenv(variants[0])
r = mepol_b200.knn_device(x, k1, algo=_lib.KNN_TC)
ws, out = r._workspace, (r.indices, r.distances, r.rk, r.stats)
base = (r.indices.clone(), r.distances.clone())
times = {v: [] for v in variants}
same = {v: True for v in variants}
for rd in range(rounds + 1):
    for v in variants:
        setenv(v)
        _lib.profile_enable(True)
        r = mepol_b200.knn_device(x, k1, algo=_lib.KNN_TC, workspace=ws, out=out)
        torch.cuda.synchronize()
        pr = _lib.profile_read()
        _lib.profile_enable(False)
        if rd == 0:
            same[v] = bool(torch.equal(base[0], r.indices) and torch.equal(base[1], r.distances))
        else:
            times[v] += pr
for v in variants:
    t = np.array(times[v])
    print(json.dumps({"N": n, "variant": v, "mean_ms": round(float(t.mean()), 4), "min_ms": round(float(t.min()), 4),
                      "std_ms": round(float(t.std()), 4), "n": int(t.size), "same_as_first": same[v],
                      "frac_sustained": round(2.0 * n * n * d / (t.mean() * 1e-3) / 1e12 / 1367.7, 4)}), flush=True)
