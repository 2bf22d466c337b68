import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mepol_b200
from mepol_b200 import _lib
n, d, k1, L = 10000, 29, int(sys.argv[1]) if len(sys.argv) > 1 else 11, 500
rng = np.random.default_rng(n + d)
x = np.cumsum(0.05 * rng.standard_normal((n // L, L, d)), axis=1).reshape(-1, d).astype(np.float32)
xt = torch.from_numpy(x).cuda()
a = mepol_b200.knn_device(xt, k1, algo=_lib.KNN_SIMT)
torch.cuda.synchronize()
print("simt ok", a.stats_dict(), flush=True)
b = mepol_b200.knn_device(xt, k1, algo=_lib.KNN_TC)
torch.cuda.synchronize()
print("tc ok", b.stats_dict(), torch.equal(a.indices, b.indices), flush=True)
