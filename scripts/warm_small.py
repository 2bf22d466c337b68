"""Tuning aid: warm-up tile count against batch size (MEPOL_TC_WARM_TILES=<w> python scripts/warm_small.py)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import mepol_b200
from mepol_b200 import _lib
from tc_lowdim import timed

out = []
for n in (8192, 16384, 32768, 65536, 131072, 262144, 524288):
    x = torch.from_numpy(np.random.default_rng(0).standard_normal((n, 29)).astype(np.float32)).cuda()
    t, r = timed(x, 5, _lib.KNN_TC, reps=10 if n <= 131072 else 3)
    out.append(f"{n}:{t:.3f}")
print(os.environ.get("MEPOL_TC_WARM_TILES", "default"), " ".join(out), flush=True)
