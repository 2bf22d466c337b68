"""CPU oracle for the importance-weighted k-NN entropy / KL objective.

TEST INFRASTRUCTURE, NOT PRODUCT CODE: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.

Numpy float64 restatement of (all paths relative to /root/reference/):
  * src/policy.py:43-61            GaussianPolicy.forward / get_log_p        -> gaussian_log_p
  * src/algorithms/mepol.py:114-139 compute_importance_weights               -> importance_weights
  * src/algorithms/mepol.py:142-154 compute_entropy                          -> entropy
  * src/algorithms/mepol.py:157-174 compute_kl                               -> kl
  * torch autograd of `loss = -entropy` (mepol.py:268-281)                   -> entropy_grad_logp,
                                                                                policy_param_grads
Pinned against the unmodified reference (torch 2.11 CPU float64, imported from /root/reference in
the build container) by oracle/gen_golden.py -> tests/golden/iw_*.npz, checked in
tests/test_oracle_iw.py.  The reference itself ships no golden vectors.
"""
import numpy as np

POLICY_EPS = 1e-7  # src/utils/dtypes.py:7 -- used only inside the variance term of get_log_p


# ----------------------------------------------------------------------------- policy (a6)
def mlp_forward(params, states):
    """params = dict(hidden=[(W,b),...], mean=(W,b), log_std=v); returns (mean, cache)."""
    h = np.asarray(states, dtype=np.float64)
    acts = [h]
    pre = []
    for W, b in params["hidden"]:
        z = h @ W.T + b
        pre.append(z)
        h = np.maximum(z, 0.0)
        acts.append(h)
    Wm, bm = params["mean"]
    mean = h @ Wm.T + bm
    return mean, (acts, pre)


def gaussian_log_p(params, states, actions):
    """policy.py:43-51: sum_a -0.5*(log 2pi + 2 log_std + (a-mu)^2/(exp(log_std)+1e-7)^2)."""
    mean, _ = mlp_forward(params, states)
    log_std = params["log_std"]
    var = (np.exp(log_std) + POLICY_EPS) ** 2
    a = np.asarray(actions, dtype=np.float64)
    return np.sum(-0.5 * (np.log(2 * np.pi) + 2 * log_std + (a - mean) ** 2 / var), axis=1)


def policy_param_grads(params, states, actions, g_logp):
    """Backprop sum_j g_logp[j] * log p(a_j|s_j) to every policy parameter (manual autograd of
    policy.py:43-61 for a ReLU MLP).  Returns a dict shaped like ``params``."""
    a = np.asarray(actions, dtype=np.float64)
    g = np.asarray(g_logp, dtype=np.float64)[:, None]
    mean, (acts, pre) = mlp_forward(params, states)
    log_std = params["log_std"]
    sd = np.exp(log_std) + POLICY_EPS
    diff = a - mean
    g_mean = g * diff / sd ** 2                                   # d logp / d mean
    # d/d log_std of -0.5*(2 log_std + diff^2/sd^2) = -1 + diff^2 * exp(log_std) / sd^3
    g_log_std = np.sum(g * (-1.0 + diff ** 2 * np.exp(log_std) / sd ** 3), axis=0)
    Wm, _ = params["mean"]
    grads = {"mean": (g_mean.T @ acts[-1], g_mean.sum(0)), "log_std": g_log_std, "hidden": []}
    gh = g_mean @ Wm
    for li in range(len(params["hidden"]) - 1, -1, -1):
        W, _ = params["hidden"][li]
        gz = gh * (pre[li] > 0.0)
        grads["hidden"].insert(0, (gz.T @ acts[li], gz.sum(0)))
        gh = gz @ W
    return grads


# ----------------------------------------------------------------------------- objective (a2-a4)
def flatten_trajectories(states, actions, real_traj_lengths):
    """Row order of the particle batch (mepol.py:98-105 / :121-135): trajectory-major, the first
    ``len`` steps of each trajectory; states[t, :len] are the states the actions were taken FROM."""
    lens = np.asarray(real_traj_lengths).reshape(-1).astype(np.int64)
    s = np.concatenate([states[t, :lens[t]] for t in range(len(lens))], axis=0)
    a = np.concatenate([actions[t, :lens[t]] for t in range(len(lens))], axis=0)
    return s, a, lens


def importance_weights(logp_target, logp_behav, traj_lengths):
    """mepol.py:121-138: per-trajectory exp(cumsum(lt - lb)), concatenated, self-normalised."""
    lens = np.asarray(traj_lengths).reshape(-1).astype(np.int64)
    d = np.asarray(logp_target, np.float64) - np.asarray(logp_behav, np.float64)
    u = np.empty_like(d)
    o = 0
    for L in lens:
        u[o:o + L] = np.exp(np.cumsum(d[o:o + L]))
        o += L
    return u / u.sum()


def _volumes(dist_k, ns, G):
    return (np.power(dist_k, ns) * np.power(np.pi, ns / 2.0)) / G      # mepol.py:150


def entropy(logp_target, logp_behav, traj_lengths, distances, indices, k, G, B, ns, eps):
    """mepol.py:142-154.  distances/indices are the [N, k+1] search outputs."""
    w = importance_weights(logp_target, logp_behav, traj_lengths)
    W = w[np.asarray(indices)[:, :-1]].sum(1)                          # :148 (self + k-1 nearest)
    V = _volumes(np.asarray(distances, np.float64)[:, k], ns, G)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        return float(-np.sum((W / k) * np.log(W / (V + eps) + eps)) + B)


def kl(logp_target, logp_behav, traj_lengths, indices, k, eps):
    """mepol.py:157-174 -> (kl clamped at 0, numeric_error flag computed before the clamp)."""
    w = importance_weights(logp_target, logp_behav, traj_lengths)
    W = w[np.asarray(indices)[:, :-1]].sum(1)
    N = w.shape[0]
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        val = (1.0 / N) * np.sum(np.log(k / (N * W) + eps))
    numeric_error = bool(np.isinf(val) or np.isnan(val))
    return float(max(0.0, val)) if not np.isnan(val) else float("nan"), numeric_error


def entropy_grad_logp(logp_target, logp_behav, traj_lengths, distances, indices, k, G, ns, eps):
    """Closed-form d(entropy)/d(logp_target[j]) (verified against the
    reference's autograd in tests/test_oracle_iw.py).  The loss gradient is the negative."""
    lens = np.asarray(traj_lengths).reshape(-1).astype(np.int64)
    idx = np.asarray(indices)[:, :-1]
    w = importance_weights(logp_target, logp_behav, lens)
    W = w[idx].sum(1)
    V = _volumes(np.asarray(distances, np.float64)[:, k], ns, G)
    q = W / (V + eps)
    gW = -(1.0 / k) * (np.log(q + eps) + q / (q + eps))
    gw = np.zeros_like(w)
    np.add.at(gw, idx.reshape(-1), np.repeat(gW, idx.shape[1]))
    s = np.dot(gw, w)
    gc = w * (gw - s)
    g = np.empty_like(gc)
    o = 0
    for L in lens:
        g[o:o + L] = np.cumsum(gc[o:o + L][::-1])[::-1]
        o += L
    return g


def params_from_state_dict(sd):
    """Build the oracle's params dict from a GaussianPolicy.state_dict() (policy.py:17-29)."""
    def arr(v):
        return np.asarray(v.detach().cpu().numpy() if hasattr(v, "detach") else v, dtype=np.float64)
    hidden = []
    i = 0
    while f"net.{i}.weight" in sd:
        hidden.append((arr(sd[f"net.{i}.weight"]), arr(sd[f"net.{i}.bias"])))
        i += 2
    return {"hidden": hidden, "mean": (arr(sd["mean.weight"]), arr(sd["mean.bias"])),
            "log_std": arr(sd["log_std"])}
