"""GPU parity on the reference's shipped configurations (BASELINE.json configs[0..3]) at their real
batch sizes, through the reference call surface: search (collect_particles_and_compute_knn's
replacement for mepol.py:190-192) + compute_entropy + backward + compute_kl, against the CPU oracle.
Rollouts cannot run here (no gym/MuJoCo): particles are synthetic with each config's shapes
(SURVEY.md section 8d2)."""
import math

import numpy as np
import pytest
import scipy.special
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu

import mepol_b200
from oracle import iw_oracle as O
from oracle import knn_oracle

#            name          T   L     F   A  hidden      ns  k   eps     particles
CONFIGS = [("grid_world",  20, 1200, 2,  2, [300, 300], 2,  50, 0.0,   "uniform6"),
           ("grid_world_k4", 20, 1200, 2, 2, [300, 300], 2, 4,  0.0,   "uniform6"),
           ("mountain_car", 20, 400, 2,  1, [300, 300], 2,  4,  1e-15, "mountaincar"),
           ("ant",          20, 500, 29, 8, [400, 300], 7,  4,  0.0,   "normal"),
           ("humanoid",     20, 500, 47, 17, [400, 300], 24, 4, 0.0,   "normal")]


class _Policy(nn.Module):
    """Same parameter layout as the reference's GaussianPolicy (src/policy.py:17-29)."""

    def __init__(self, hidden, F, A):
        super().__init__()
        self.log_std = nn.Parameter(-0.5 * torch.ones(A, dtype=torch.float64))
        layers, prev = [], F
        for h in hidden:
            layers += [nn.Linear(prev, h).double(), nn.ReLU()]
            prev = h
        self.net = nn.Sequential(*layers)
        self.mean = nn.Linear(prev, A).double()


@pytest.mark.parametrize("cfg", CONFIGS, ids=[c[0] for c in CONFIGS])
def test_config_shapes_match_oracle(cfg):
    name, T, L, F, A, hidden, ns, k, eps, kind = cfg
    rng = np.random.default_rng(abs(hash(name)) % 2**31)
    torch.manual_seed(0)
    states = rng.standard_normal((T, L + 1, F)).astype(np.float32)
    if kind == "uniform6":
        states[..., :2] = rng.uniform(-6, 6, (T, L + 1, 2))
    elif kind == "mountaincar":
        states[..., 0] = rng.uniform(-1.2, 0.6, (T, L + 1))
        states[..., 1] = rng.uniform(-0.07, 0.07, (T, L + 1))
    actions = (0.6 * rng.standard_normal((T, L, A))).astype(np.float32)
    lens = np.full((T, 1), L, dtype=np.int64)
    next_states = np.ascontiguousarray(states[:, 1:, :ns].reshape(-1, ns))

    # --- search (replacement of mepol.py:190-192) vs the oracle
    dist, idx, handle = mepol_b200.search_particles(next_states, k)
    ref_d, ref_i = knn_oracle.knn_c(next_states, k + 1, mode="kd_tree")
    np.testing.assert_array_equal(idx.numpy(), ref_i)
    np.testing.assert_allclose(dist.numpy(), ref_d, rtol=1e-12)

    # --- objective through the reference signatures vs the oracle
    beh, tgt = _Policy(hidden, F, A), _Policy(hidden, F, A)
    tgt.load_state_dict(beh.state_dict())
    with torch.no_grad():
        for p in tgt.parameters():
            p.add_(2e-3 * torch.randn_like(p))
    G = float(scipy.special.gamma(ns / 2 + 1))
    B = float(np.log(k) - scipy.special.digamma(k))
    t_states, t_actions = torch.tensor(states, dtype=torch.float64), torch.tensor(actions, dtype=torch.float64)
    t_lens = torch.tensor(lens)
    ent = mepol_b200.compute_entropy(beh, tgt, t_states, t_actions, T, t_lens, dist, idx, k, G, B, ns, eps)
    (-ent).backward()
    with torch.no_grad():
        kl, kl_err = mepol_b200.compute_kl(beh, tgt, t_states, t_actions, T, t_lens, dist, idx, k, eps)

    s, a, ln = O.flatten_trajectories(states.astype(np.float64), actions.astype(np.float64), lens)
    pt, pb = O.params_from_state_dict(tgt.state_dict()), O.params_from_state_dict(beh.state_dict())
    lt, lb = O.gaussian_log_p(pt, s, a), O.gaussian_log_p(pb, s, a)
    H_ref = O.entropy(lt, lb, ln, ref_d, ref_i, k, G, B, ns, eps)
    kl_ref, err_ref = O.kl(lt, lb, ln, ref_i, k, eps)
    g_ref = O.policy_param_grads(pt, s, a, -O.entropy_grad_logp(lt, lb, ln, ref_d, ref_i, k, G, ns, eps))
    assert math.isfinite(H_ref)
    np.testing.assert_allclose(ent.item(), H_ref, rtol=1e-9)
    np.testing.assert_allclose(kl.item(), kl_ref, rtol=1e-7, atol=1e-12)
    assert bool(kl_err) == err_ref
    named = dict(tgt.named_parameters())
    for i, (gw, gb) in enumerate(g_ref["hidden"]):
        for suffix, val in (("weight", gw), ("bias", gb)):
            got = named[f"net.{2 * i}.{suffix}"].grad.numpy()
            sc = np.abs(val).max() + 1e-300
            np.testing.assert_allclose(got / sc, val / sc, rtol=0, atol=1e-8)
    for suffix, val in (("weight", g_ref["mean"][0]), ("bias", g_ref["mean"][1])):
        got = named[f"mean.{suffix}"].grad.numpy()
        sc = np.abs(val).max() + 1e-300
        np.testing.assert_allclose(got / sc, val / sc, rtol=0, atol=1e-8)
    sc = np.abs(g_ref["log_std"]).max() + 1e-300
    np.testing.assert_allclose(named["log_std"].grad.numpy() / sc, g_ref["log_std"] / sc, rtol=0, atol=1e-8)


FULL = [("ant_full", 100, 500, 29, 8, [400, 300], 7, 4, 0.0, "normal"),
        ("grid_world_full", 100, 1200, 2, 2, [300, 300], 2, 50, 0.0, "uniform6"),
        ("mountain_car_full", 100, 400, 2, 1, [300, 300], 2, 4, 1e-15, "mountaincar")]


@pytest.mark.parametrize("cfg", FULL, ids=[c[0] for c in FULL])
def test_full_entropy_evaluation_call_shape(cfg):
    """The full-entropy evaluation of the reference loop (mepol.py:347-354 and :512-519): 5x the batch
    (num_traj * full_entropy_traj_scale), `full_entropy_k`, the SAME policy object passed as behavioural
    and target, under no_grad.  The weights are then exactly 1/N; checked against the oracle, against a
    second call (the behavioural cache must not confuse the two roles), and the result carries no graph."""
    name, T, L, F, A, hidden, ns, k, eps, kind = cfg
    rng = np.random.default_rng(abs(hash(name)) % 2**31)
    states = rng.standard_normal((T, L + 1, F)).astype(np.float32)
    if kind == "uniform6":
        states[..., :2] = rng.uniform(-6, 6, (T, L + 1, 2))
    elif kind == "mountaincar":
        states[..., 0] = rng.uniform(-1.2, 0.6, (T, L + 1))
        states[..., 1] = rng.uniform(-0.07, 0.07, (T, L + 1))
    actions = (0.6 * rng.standard_normal((T, L, A))).astype(np.float32)
    lens = np.full((T, 1), L, dtype=np.int64)
    next_states = np.ascontiguousarray(states[:, 1:, :ns].reshape(-1, ns))
    n = T * L
    dist, idx, _ = mepol_b200.search_particles(next_states, k)
    rows = np.sort(rng.choice(n, 512, replace=False))
    ref_d, ref_i = knn_oracle.knn_c(next_states, k + 1, queries=next_states[rows], mode="kd_tree")
    np.testing.assert_array_equal(idx.numpy()[rows], ref_i)
    np.testing.assert_allclose(dist.numpy()[rows], ref_d, rtol=1e-12)

    torch.manual_seed(1)
    pol = _Policy(hidden, F, A)
    G = float(scipy.special.gamma(ns / 2 + 1))
    full_B = float(np.log(k) - scipy.special.digamma(k))
    t_states, t_actions = torch.tensor(states, dtype=torch.float64), torch.tensor(actions, dtype=torch.float64)
    t_lens = torch.tensor(lens)
    state0 = torch.get_rng_state()
    with torch.no_grad():
        ent = mepol_b200.compute_entropy(pol, pol, t_states, t_actions, T, t_lens, dist, idx, k, G, full_B, ns, eps)
        ent2 = mepol_b200.compute_entropy(pol, pol, t_states, t_actions, T, t_lens, dist, idx, k, G, full_B, ns, eps)
    assert ent.device.type == "cpu" and ent.dtype == torch.float64 and not ent.requires_grad
    assert torch.equal(ent, ent2)
    # quirk Q7: two get_log_p calls per trajectory burn randn(L, A) each from the global generator
    torch.set_rng_state(state0)
    for _ in range(2):
        for _t in range(T):
            for _c in range(2):
                torch.randn((L, A), dtype=torch.float64)
    expect_state = torch.get_rng_state()
    torch.set_rng_state(state0)
    with torch.no_grad():
        mepol_b200.compute_entropy(pol, pol, t_states, t_actions, T, t_lens, dist, idx, k, G, full_B, ns, eps)
        mepol_b200.compute_entropy(pol, pol, t_states, t_actions, T, t_lens, dist, idx, k, G, full_B, ns, eps)
    assert torch.equal(torch.get_rng_state(), expect_state)
    zeros = np.zeros(n)
    H_ref = O.entropy(zeros, zeros, lens.reshape(-1), dist.numpy(), idx.numpy(), k, G, full_B, ns, eps)
    assert math.isfinite(H_ref)
    np.testing.assert_allclose(ent.item(), H_ref, rtol=1e-10)
