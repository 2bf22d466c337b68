"""Timings of the reference's shipped configurations (BASELINE.json configs[0..3]) on the GPU box:
search (sklearn on the host cores vs mepol_b200 through the host-buffer API) and one
policy_update + compute_kl iteration through the reference call surface (GPU drop-in vs a CPU
torch float64 restatement of the same call sequence on the same policies)."""
import json
import os
import sys
import time

import numpy as np
import scipy.special
import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mepol_b200
from oracle import knn_oracle

CONFIGS = [("grid_world", 20, 1200, 2, 2, [300, 300], 2, 50, 0.0), ("mountain_car", 20, 400, 2, 1, [300, 300], 2, 4, 1e-15),
           ("ant", 20, 500, 29, 8, [400, 300], 7, 4, 0.0), ("humanoid", 20, 500, 47, 17, [400, 300], 24, 4, 0.0),
           ("ant_full_entropy_5x", 100, 500, 29, 8, [400, 300], 7, 4, 0.0)]


class Policy(nn.Module):
    def __init__(self, hidden, F, A):
        super().__init__()
        self.log_std = nn.Parameter(-0.5 * torch.ones(A, dtype=torch.float64))
        layers, prev = [], F
        for h in hidden:
            layers += [nn.Linear(prev, h).double(), nn.ReLU()]
            prev = h
        self.net = nn.Sequential(*layers)
        self.mean = nn.Linear(prev, A).double()

    def get_log_p(self, s, a):
        mean = self.mean(self.net(s))
        return torch.sum(-0.5 * (np.log(2 * np.pi) + 2 * self.log_std + (a - mean) ** 2 / (torch.exp(self.log_std) + 1e-7) ** 2), dim=1)


def cpu_entropy(beh, tgt, states, actions, T, lens, dist, idx, k, G, B, ns, eps):
    """CPU torch float64 restatement of compute_entropy (mepol.py:114-154) for timing only."""
    parts = []
    for t in range(T):
        L = int(lens[t, 0])
        d = tgt.get_log_p(states[t, :L], actions[t, :L]) - beh.get_log_p(states[t, :L], actions[t, :L])
        parts.append(torch.exp(torch.cumsum(d, 0)))
    w = torch.cat(parts)
    w = w / w.sum()
    W = torch.sum(w[idx[:, :-1]], dim=1)
    V = (torch.pow(dist[:, k], ns) * np.pi ** (ns / 2)) / G
    return -torch.sum((W / k) * torch.log(W / (V + eps) + eps)) + B


def main():
    rows = []
    for name, T, L, F, A, hidden, ns, k, eps in CONFIGS:
        rng = np.random.default_rng(1)
        torch.manual_seed(0)
        states = rng.standard_normal((T, L + 1, F)).astype(np.float32)
        actions = (0.6 * rng.standard_normal((T, L, A))).astype(np.float32)
        nxt = np.ascontiguousarray(states[:, 1:, :ns].reshape(-1, ns))
        n = nxt.shape[0]
        t0 = time.perf_counter(); knn_oracle.knn_sklearn(nxt, k + 1, n_jobs=1); t_sk = time.perf_counter() - t0
        mepol_b200.search_particles(nxt, k)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            dist, idx, _h = mepol_b200.search_particles(nxt, k)
        torch.cuda.synchronize(); t_gpu = (time.perf_counter() - t0) / 5
        beh, tgt = Policy(hidden, F, A), Policy(hidden, F, A)
        tgt.load_state_dict(beh.state_dict())
        opt = torch.optim.Adam(tgt.parameters(), lr=1e-5)
        G = float(scipy.special.gamma(ns / 2 + 1)); B = float(np.log(k) - scipy.special.digamma(k))
        ts, ta = torch.tensor(states, dtype=torch.float64), torch.tensor(actions, dtype=torch.float64)
        tl = torch.full((T, 1), L, dtype=torch.int64)
        args = (ts, ta, T, tl, dist, idx, k, G, B, ns, eps)
        mepol_b200.policy_update(opt, beh, tgt, *args)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            mepol_b200.policy_update(opt, beh, tgt, *args)
            with torch.no_grad():
                mepol_b200.compute_kl(beh, tgt, ts, ta, T, tl, dist, idx, k, eps)
        torch.cuda.synchronize(); t_step_gpu = (time.perf_counter() - t0) / 5
        t0 = time.perf_counter()
        for _ in range(2):
            opt.zero_grad()
            loss = -cpu_entropy(beh, tgt, ts, ta, T, tl, dist, idx, k, G, B, ns, eps)
            loss.backward(); opt.step()
            with torch.no_grad():
                cpu_entropy(beh, tgt, ts, ta, T, tl, dist, idx, k, G, B, ns, eps)   # stands in for compute_kl's second IW evaluation
        t_step_cpu = (time.perf_counter() - t0) / 2
        rows.append(dict(config=name, N=n, ns=ns, k=k, sklearn_search_s=t_sk, gpu_search_s=t_gpu, cpu_step_s=t_step_cpu, gpu_step_s=t_step_gpu))
        print(json.dumps(rows[-1]), flush=True)
    print("cores", os.cpu_count(), "torch threads", torch.get_num_threads())


if __name__ == "__main__":
    main()
