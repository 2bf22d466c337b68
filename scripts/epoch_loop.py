"""Epoch-realistic timing of the drop-in (mepol.py:404-470): every epoch brings a NEW batch of
particles (new search, new cached batch, new captured graphs), followed by max_off_iters = 30
off-policy steps (policy_update + compute_kl).  Compares CUDA-graph replay with eager launches."""
import os
import sys
import time

import numpy as np
import scipy.special
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import mepol_b200
from config_timings import Policy

CFG = {"ant": (20, 500, 29, 8, [400, 300], 7, 4), "humanoid": (20, 500, 47, 17, [400, 300], 24, 4),
       "grid_world": (20, 1200, 2, 2, [300, 300], 2, 50), "mountain_car": (20, 400, 2, 1, [300, 300], 2, 4)}
for name, (T_, L, F, A, hidden, ns, k) in CFG.items():
    for graphs in (True, False):
        mepol_b200.configure(cuda_graphs=graphs)
        rng = np.random.default_rng(1)
        torch.manual_seed(0)
        beh, tgt = Policy(hidden, F, A), Policy(hidden, F, A)
        tgt.load_state_dict(beh.state_dict())
        opt = torch.optim.Adam(tgt.parameters(), lr=1e-5)
        G = float(scipy.special.gamma(ns / 2 + 1))
        B = float(np.log(k) - scipy.special.digamma(k))
        t_search = t_steps = 0.0
        n_epochs = 6
        for epoch in range(n_epochs):
            states = rng.standard_normal((T_, L + 1, F)).astype(np.float32)
            actions = (0.6 * rng.standard_normal((T_, L, A))).astype(np.float32)
            nxt = np.ascontiguousarray(states[:, 1:, :ns].reshape(-1, ns))
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            dist, idx, _h = mepol_b200.search_particles(nxt, k)
            ts, ta = torch.tensor(states, dtype=torch.float64), torch.tensor(actions, dtype=torch.float64)
            tl = torch.full((T_, 1), L, dtype=torch.int64)
            t1 = time.perf_counter()
            for it in range(30):
                mepol_b200.policy_update(opt, beh, tgt, ts, ta, T_, tl, dist, idx, k, G, B, ns, 0.0)
                with torch.no_grad():
                    mepol_b200.compute_kl(beh, tgt, ts, ta, T_, tl, dist, idx, k, 0.0)
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            if epoch > 0:           # epoch 0 pays library / allocator warm-up
                t_search += t1 - t0
                t_steps += t2 - t1
            beh.load_state_dict(tgt.state_dict())
        print(f"{name:13s} graphs={graphs!s:5s}  search+staging {1e3 * t_search / (n_epochs - 1):7.2f} ms/epoch   "
              f"30 steps {1e3 * t_steps / (n_epochs - 1):7.2f} ms/epoch  ({1e3 * t_steps / (n_epochs - 1) / 30:.2f} ms/step)", flush=True)
mepol_b200.configure(cuda_graphs=True)
