"""Generate tests/golden/*.npz from the UNMODIFIED reference, run in the build container only.

    python oracle/gen_golden.py            # needs /root/reference (absent on the GPU box)

TEST INFRASTRUCTURE.  The reference ships no golden vectors, so parity is
pinned on outputs of the reference itself executed here:
  * k-NN:  sklearn.neighbors.NearestNeighbors exactly as /root/reference/src/algorithms/mepol.py:190-192
  * objective: compute_entropy / compute_kl / policy_update-style backward imported from
    /root/reference/src/algorithms/mepol.py with GaussianPolicy from /root/reference/src/policy.py
    (matplotlib is stubbed: only get_heatmap uses it).
Every file records library versions and the seed.
"""
import os
import sys
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def import_reference():
    for name in ("matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.use = lambda *a, **k: None
            sys.modules[name] = m
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import warnings
    warnings.filterwarnings("ignore")
    import src.algorithms.mepol as ref_mepol
    import src.policy as ref_policy
    return ref_mepol, ref_policy


def versions():
    import sklearn, torch, scipy
    return np.array([f"sklearn={sklearn.__version__}", f"torch={torch.__version__}",
                     f"numpy={np.__version__}", f"scipy={scipy.__version__}"])


def knn_cases():
    # name, N, d, k1, kind
    return [
        ("d29_k4", 1536, 29, 5, "normal"),
        ("d24_k4", 1024, 24, 5, "normal"),
        ("d7_k4", 2048, 7, 5, "normal"),
        ("d3_k4", 1024, 3, 5, "normal"),
        ("d2_k50", 1500, 2, 51, "uniform"),
        ("d2_k4_mc", 1200, 2, 5, "mountaincar"),
        ("d29_k1", 700, 29, 2, "normal"),
        ("d29_k4_offset", 1024, 29, 5, "offset"),
        ("d29_k4_dups", 600, 29, 5, "dups"),
        ("d2_k4_dups", 500, 2, 5, "dups"),
        ("d47_k4", 640, 47, 5, "normal"),
        ("d29_k4_traj", 1500, 29, 5, "traj"),
    ]


def make_x(n, d, kind, rng):
    if kind == "normal":
        x = rng.standard_normal((n, d))
    elif kind == "uniform":
        x = rng.uniform(-6, 6, (n, d))
    elif kind == "mountaincar":
        x = np.stack([rng.uniform(-1.2, 0.6, n), rng.uniform(-0.07, 0.07, n)], 1)
    elif kind == "offset":
        x = rng.standard_normal((n, d)) * 0.3 + 40.0 + np.arange(d)[None, :]
    elif kind == "traj":
        x = np.cumsum(0.05 * rng.standard_normal((n // 100, 100, d)), axis=1).reshape(-1, d)
    elif kind == "dups":
        x = rng.standard_normal((n, d))
        x[100:108] = x[100]          # 8 exact duplicates (Q2 "knn aliasing")
        x[300:303] = x[300]          # a triple
        x[-1] = x[0]                 # a far-apart pair
    return np.ascontiguousarray(x, dtype=np.float32)


def gen_knn():
    from sklearn.neighbors import NearestNeighbors
    for seed, (name, n, d, k1, kind) in enumerate(knn_cases()):
        rng = np.random.default_rng(1000 + seed)
        x = make_x(n, d, kind, rng)
        nbrs = NearestNeighbors(n_neighbors=k1, metric="euclidean", algorithm="auto", n_jobs=1)
        nbrs.fit(x)
        dist, idx = nbrs.kneighbors(x)
        np.savez_compressed(os.path.join(OUT, f"knn_{name}.npz"), x=x, dist=dist, idx=idx.astype(np.int32),
                            k1=k1, method=nbrs._fit_method, seed=1000 + seed, kind=kind,
                            versions=versions())
        print(f"knn_{name}: N={n} d={d} k1={k1} method={nbrs._fit_method}")


def iw_cases():
    # name, T, L, F, A, hidden, ns, k, eps, ragged, perturb
    return [
        ("ant_like", 6, 50, 29, 8, [32, 24], 7, 4, 0.0, False, 1e-2),
        ("ragged", 5, 40, 11, 3, [16, 16], 11, 4, 0.0, True, 3e-2),
        ("mc_like_eps", 4, 60, 2, 1, [16, 16], 2, 4, 1e-15, False, 2e-2),
        ("grid_like_k50", 4, 80, 2, 2, [16, 16], 2, 50, 0.0, False, 1e-2),
        ("same_policy", 4, 30, 29, 8, [16, 8], 29, 4, 0.0, False, 0.0),
        ("humanoid_like", 5, 40, 47, 17, [24, 16], 24, 4, 0.0, False, 5e-3),
    ]


def gen_iw():
    import scipy.special
    import torch
    ref_mepol, ref_policy = import_reference()
    from sklearn.neighbors import NearestNeighbors
    for ci, (name, T, L, F, A, hidden, ns, k, eps, ragged, perturb) in enumerate(iw_cases()):
        seed = 2000 + ci
        rng = np.random.default_rng(seed)
        torch.manual_seed(seed)
        states = rng.standard_normal((T, L + 1, F)).astype(np.float32)
        actions = (0.6 * rng.standard_normal((T, L, A))).astype(np.float32)
        lens = np.full((T, 1), L, dtype=np.int32)
        if ragged:
            lens[:, 0] = rng.integers(L // 3, L + 1, T)
            lens[0, 0] = L
        next_states = np.concatenate([states[t, 1:lens[t, 0] + 1, :ns] for t in range(T)], 0)
        nbrs = NearestNeighbors(n_neighbors=k + 1, metric="euclidean", algorithm="auto", n_jobs=1)
        nbrs.fit(next_states)
        dist, idx = nbrs.kneighbors(next_states)

        beh = ref_policy.GaussianPolicy(hidden, F, A)
        tgt = ref_policy.GaussianPolicy(hidden, F, A)
        tgt.load_state_dict(beh.state_dict())
        with torch.no_grad():
            for p in tgt.parameters():
                p.add_(perturb * torch.randn_like(p))
        G = scipy.special.gamma(ns / 2 + 1)
        B = np.log(k) - scipy.special.digamma(k)

        t_states = torch.tensor(states, dtype=torch.float64)
        t_actions = torch.tensor(actions, dtype=torch.float64)
        t_lens = torch.tensor(lens, dtype=torch.int64)
        t_dist = torch.tensor(dist, dtype=torch.float64)
        t_idx = torch.tensor(idx, dtype=torch.int64)

        ent = ref_mepol.compute_entropy(beh, tgt, t_states, t_actions, T, t_lens, t_dist, t_idx,
                                        k, G, B, ns, eps)
        loss = -ent
        for p in tgt.parameters():
            p.grad = None
        loss.backward()
        grads = {n_: p.grad.detach().numpy().copy() for n_, p in tgt.named_parameters()}
        with torch.no_grad():
            klv, kl_err = ref_mepol.compute_kl(beh, tgt, t_states, t_actions, T, t_lens, t_dist,
                                               t_idx, k, eps)
            w = ref_mepol.compute_importance_weights(beh, tgt, t_states, t_actions, T, t_lens)
            s_flat = torch.cat([t_states[t, :lens[t, 0]] for t in range(T)], 0)
            a_flat = torch.cat([t_actions[t, :lens[t, 0]] for t in range(T)], 0)
            lt = tgt.get_log_p(s_flat, a_flat).numpy()
            lb = beh.get_log_p(s_flat, a_flat).numpy()
        # d(loss)/d(logp_target) through the reference's own objective, log-probs as leaves
        lt_leaf = torch.tensor(lt, dtype=torch.float64, requires_grad=True)
        lb_t = torch.tensor(lb, dtype=torch.float64)
        parts, o = [], 0
        for t in range(T):
            Lt = int(lens[t, 0])
            parts.append(torch.exp(torch.cumsum(lt_leaf[o:o + Lt] - lb_t[o:o + Lt], dim=0)))
            o += Lt
        wl = torch.cat(parts)
        wl = wl / wl.sum()
        Wl = torch.sum(wl[t_idx[:, :-1]], dim=1)
        vol = (torch.pow(t_dist[:, k], ns) * torch.pow(torch.tensor(np.pi), ns / 2)) / G
        ent_l = -torch.sum((Wl / k) * torch.log((Wl / (vol + eps)) + eps)) + B
        (-ent_l).backward()

        payload = dict(
            states=states, actions=actions, lens=lens, dist=dist, idx=idx.astype(np.int32),
            hidden=np.array(hidden), F=F, A=A, ns=ns, k=k, eps=eps, G=G, B=B, seed=seed,
            entropy=ent.detach().numpy(), kl=klv.numpy(), kl_numeric_error=bool(kl_err),
            w=w.numpy(), logp_target=lt, logp_behav=lb, grad_loss_logp=lt_leaf.grad.numpy(),
            versions=versions())
        for n_, p in beh.state_dict().items():
            payload["beh." + n_] = p.numpy()
        for n_, p in tgt.state_dict().items():
            payload["tgt." + n_] = p.numpy()
        for n_, g in grads.items():
            payload["grad." + n_] = g
        np.savez_compressed(os.path.join(OUT, f"iw_{name}.npz"), **payload)
        print(f"iw_{name}: N={next_states.shape[0]} H={float(ent):.6f} KL={float(klv):.3e} err={bool(kl_err)}")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    gen_knn()
    gen_iw()
