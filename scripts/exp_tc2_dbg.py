"""Timing experiments on the shipped pruning kernel (MEPOL_NVCC_EXTRA=-DMEPOL_TC2_DBGMODES builds only).
MEPOL_TC2_DBG: 1 = accumulator 1 idle, 2 = no TMEM load, 4 = load + release only, 16 = load + min tree,
32 = load + tree + vote (no survivor path), 33 / 5 = 32 / 4 with accumulator 1 idle; MEPOL_TC2_KSTEPS = MMAs per tile.
Results of these modes are invalid (the library stops after the pruning kernel); only the kernel time is read.
Modes alternate launch by launch (power-cap drift), means are reported.
    python scripts/exp_tc2_dbg.py [N] [rounds]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mepol_b200
from mepol_b200 import _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1048576
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 6
d, k1 = 29, 5
x = torch.from_numpy(np.random.default_rng(0).standard_normal((n, d), dtype=np.float32)).cuda()
r = mepol_b200.knn_device(x, k1, algo=_lib.KNN_TC)
ws, out = r._workspace, (r.indices, r.distances, r.rk, r.stats)
modes = [tuple(int(v) for v in m.split(':')) for m in os.environ.get('EXP_MODES', '0:2,32:2,16:2,4:2,2:2,1:2,33:2,5:2,0:1,32:1,4:1').split(',')]
times = {m: [] for m in modes}
for rd in range(rounds + 1):
    for dbg, ks in modes:
        os.environ["MEPOL_TC2_DBG"] = str(dbg)
        os.environ["MEPOL_TC2_KSTEPS"] = str(ks)
        _lib.profile_enable(True)
        mepol_b200.knn_device(x, k1, algo=_lib.KNN_TC, workspace=ws, out=out)
        torch.cuda.synchronize()
        pr = _lib.profile_read()
        _lib.profile_enable(False)
        if rd:
            times[(dbg, ks)] += pr
for (dbg, ks), t in times.items():
    t = np.array(t)
    cyc = t.mean() * 1e-3 * 1.9e9 / ((n / 256) * (n / 256) / 148)
    print(json.dumps({"N": n, "dbg": dbg, "ksteps": ks, "mean_ms": round(float(t.mean()), 3), "min_ms": round(float(t.min()), 3),
                      "std_ms": round(float(t.std()), 3), "cycles_per_cta_tile_at_1.9GHz": round(cyc)}), flush=True)
