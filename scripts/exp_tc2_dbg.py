"""Timing experiments on the shipped pruning kernel (bring-up build only: MEPOL_NVCC_EXTRA=-DMEPOL_TC2_DBGMODES).
MEPOL_TC2_DBG bits: 1 = accumulator 1 idle, 2 = no TMEM load, 4 = load + release only, 8 = A operand read from TMEM,
16 = load + min tree only, 32 = load + tree + vote only; MEPOL_TC2_KSTEPS = MMAs per tile.
Results of these modes are invalid; only the kernel time is read.
    python scripts/exp_tc2_dbg.py [N]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mepol_b200
from mepol_b200 import _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 524288
d, k1 = 29, 5
x = torch.from_numpy(np.random.default_rng(0).standard_normal((n, d), dtype=np.float32)).cuda()
r = mepol_b200.knn_device(x, k1, algo=_lib.KNN_TC)
ws, out = r._workspace, (r.indices, r.distances, r.rk, r.stats)
for dbg, ks in [(0, 2), (4, 2), (16, 2), (32, 2), (12, 2), (8, 2), (24, 2), (4, 1), (16, 1), (0, 1), (0, 2), (4, 2), (16, 2), (32, 2), (12, 2)]:
    os.environ["MEPOL_TC2_DBG"] = str(dbg)
    os.environ["MEPOL_TC2_KSTEPS"] = str(ks)
    ts = []
    for _ in range(3):
        _lib.profile_enable(True)
        try:
            mepol_b200.knn_device(x, k1, algo=_lib.KNN_TC, workspace=ws, out=out)
        except Exception as e:          # invalid scores may overflow candidate buffers etc.; the prune time is still recorded
            pass
        torch.cuda.synchronize()
        pr = _lib.profile_read()
        _lib.profile_enable(False)
        ts += pr
    cyc = min(ts) * 1e-3 * 1.9e9 / ((n / 256) * (n / 256) / 148)
    print(json.dumps({"N": n, "dbg": dbg, "ksteps": ks, "prune_ms": [round(t, 3) for t in ts], "cycles_per_cta_tile_at_1.9GHz": round(cyc)}), flush=True)
