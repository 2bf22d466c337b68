import os, sys
import numpy as np, torch
sys.path.insert(0, "/root/repo")
import mepol_b200
from mepol_b200 import _lib
for n in (1048576, 131072, 65536):
    x = torch.from_numpy(np.random.default_rng(0).standard_normal((n, 29), dtype=np.float32)).cuda()
    r = mepol_b200.knn_device(x, 5, algo=_lib.KNN_TC)
    torch.cuda.synchronize()
    for single in ("",):
        if single: os.environ["MEPOL_TC2_SINGLE"] = "1"
        else: os.environ.pop("MEPOL_TC2_SINGLE", None)
        os.environ["MEPOL_TC_PROF"] = "1"
        sys.stderr.write(f"--- N={n} single={single!r}\n"); sys.stderr.flush()
        r = mepol_b200.knn_device(x, 5, algo=_lib.KNN_TC, workspace=r._workspace, out=(r.indices, r.distances, r.rk, r.stats))
        torch.cuda.synchronize()
        os.environ.pop("MEPOL_TC_PROF")
