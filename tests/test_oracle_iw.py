"""The objective oracle (oracle/iw_oracle.py) is pinned against outputs of the unmodified reference
(compute_entropy / compute_kl / autograd) stored in tests/golden/iw_*.npz."""
import numpy as np
import pytest

from conftest import golden_files, load_golden
from oracle import iw_oracle as O


def _params(g, who):
    sd = {k[len(who) + 1:]: v for k, v in g.items() if k.startswith(who + ".")}
    return O.params_from_state_dict(sd)


@pytest.mark.parametrize("path", golden_files("iw_"), ids=lambda p: p.split("iw_")[-1][:-4])
def test_oracle_matches_reference_golden(path):
    g = load_golden(path)
    s, a, lens = O.flatten_trajectories(g["states"], g["actions"], g["lens"])
    pt, pb = _params(g, "tgt"), _params(g, "beh")
    lt = O.gaussian_log_p(pt, s, a)
    lb = O.gaussian_log_p(pb, s, a)
    np.testing.assert_allclose(lt, g["logp_target"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(lb, g["logp_behav"], rtol=1e-12, atol=1e-12)
    k, ns, eps = int(g["k"]), int(g["ns"]), float(g["eps"])
    G, B = float(g["G"]), float(g["B"])
    w = O.importance_weights(lt, lb, lens)
    np.testing.assert_allclose(w, g["w"], rtol=1e-10)
    H = O.entropy(lt, lb, lens, g["dist"], g["idx"], k, G, B, ns, eps)
    np.testing.assert_allclose(H, float(g["entropy"]), rtol=1e-11)
    klv, err = O.kl(lt, lb, lens, g["idx"], k, eps)
    assert err == bool(g["kl_numeric_error"])
    np.testing.assert_allclose(klv, float(g["kl"]), rtol=1e-9, atol=1e-13)
    g_ent = O.entropy_grad_logp(lt, lb, lens, g["dist"], g["idx"], k, G, ns, eps)
    scale = np.max(np.abs(g["grad_loss_logp"])) + 1e-300
    np.testing.assert_allclose(-g_ent / scale, g["grad_loss_logp"] / scale, rtol=0, atol=1e-11)
    pg = O.policy_param_grads(pt, s, a, -g_ent)
    names = [("net.%d" % (2 * i), pg["hidden"][i]) for i in range(len(pg["hidden"]))] + [("mean", pg["mean"])]
    for name, (gw, gb) in names:
        for suffix, val in (("weight", gw), ("bias", gb)):
            ref = g[f"grad.{name}.{suffix}"]
            sc = np.max(np.abs(ref)) + 1e-300
            np.testing.assert_allclose(val / sc, ref / sc, rtol=0, atol=1e-10)
    ref = g["grad.log_std"]
    np.testing.assert_allclose(pg["log_std"], ref, rtol=1e-9, atol=1e-10 * np.max(np.abs(ref)))


def test_same_policy_gives_uniform_weights():
    g = load_golden(golden_files("iw_same_policy")[0])
    n = g["w"].shape[0]
    np.testing.assert_allclose(g["w"], np.full(n, 1.0 / n), rtol=1e-12)
    lens = g["lens"].reshape(-1)
    w = O.importance_weights(g["logp_behav"], g["logp_behav"], lens)
    np.testing.assert_allclose(w, 1.0 / n, rtol=1e-14)


def test_shift_invariance_of_weights():
    rng = np.random.default_rng(0)
    lens = np.array([30, 17, 25])
    lt, lb = rng.normal(size=72), rng.normal(size=72)
    # adding a per-step constant to every log-ratio multiplies u by exp(c*t): NOT invariant;
    # but scaling all u by a constant (shift of the first step of every trajectory) is.
    w0 = O.importance_weights(lt, lb, lens)
    lt2 = lt.copy()
    off = np.concatenate([[0], np.cumsum(lens)[:-1]])
    lt2[off] += 3.0
    np.testing.assert_allclose(O.importance_weights(lt2, lb, lens), w0, rtol=1e-12)


def test_kl_numeric_error_flag():
    lens = np.array([4])
    lt = np.array([0.0, 800.0, 0.0, 0.0])     # exp overflows -> inf/inf = nan weights
    lb = np.zeros(4)
    idx = np.array([[0, 1, 2], [1, 0, 2], [2, 3, 1], [3, 2, 1]])
    _, err = O.kl(lt, lb, lens, idx, 2, 0.0)
    assert err
