"""cProfile of one compute_kl / policy_update call through the reference surface (Ant shapes)."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import scipy.special
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import mepol_b200
from config_timings import Policy

T_, L, F, A, hidden, ns, k, eps = 20, 500, 29, 8, [400, 300], 7, 4, 0.0
rng = np.random.default_rng(1)
torch.manual_seed(0)
states = rng.standard_normal((T_, L + 1, F)).astype(np.float32)
actions = (0.6 * rng.standard_normal((T_, L, A))).astype(np.float32)
nxt = np.ascontiguousarray(states[:, 1:, :ns].reshape(-1, ns))
dist, idx, _h = mepol_b200.search_particles(nxt, k)
beh, tgt = Policy(hidden, F, A), Policy(hidden, F, A)
tgt.load_state_dict(beh.state_dict())
opt = torch.optim.Adam(tgt.parameters(), lr=1e-5)
G = float(scipy.special.gamma(ns / 2 + 1))
B = float(np.log(k) - scipy.special.digamma(k))
ts, ta = torch.tensor(states, dtype=torch.float64), torch.tensor(actions, dtype=torch.float64)
tl = torch.full((T_, 1), L, dtype=torch.int64)
args = (ts, ta, T_, tl, dist, idx, k, G, B, ns, eps)
for _ in range(5):
    mepol_b200.policy_update(opt, beh, tgt, *args)
    mepol_b200.compute_kl(beh, tgt, ts, ta, T_, tl, dist, idx, k, eps)
for burn in (True, False):
    mepol_b200.entropy.configure(burn_rng=burn)
    for name, fn in (("compute_kl", lambda: mepol_b200.compute_kl(beh, tgt, ts, ta, T_, tl, dist, idx, k, eps)),
                     ("policy_update", lambda: mepol_b200.policy_update(opt, beh, tgt, *args))):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(50):
            fn()
        torch.cuda.synchronize()
        print(f"burn={burn} {name}: {(time.perf_counter() - t0) / 50 * 1e3:.3f} ms")
mepol_b200.entropy.configure(burn_rng=True)
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
    mepol_b200.compute_kl(beh, tgt, ts, ta, T_, tl, dist, idx, k, eps)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
