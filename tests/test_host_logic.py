"""Host-side logic that needs no GPU: trajectory offsets, the rollout collector's output contract
(mepol.py:70-111), the drop-in installer, and a world_size-2 gloo run of the row partitioner."""
import os
import sys
import types

import numpy as np
import pytest
import torch

import mepol_b200
from mepol_b200 import dropin


class _ToyEnv:
    """Deterministic 3-feature env with an early `done` on trajectory 1 (ragged lengths)."""
    num_features = 3

    class _Space:
        shape = (2,)
    action_space = _Space()

    def __init__(self):
        self.t = 0
        self.ep = -1

    def reset(self):
        self.ep += 1
        self.t = 0
        return np.array([self.ep, 0.0, 0.5], dtype=np.float64)

    def step(self, a):
        self.t += 1
        done = self.ep == 1 and self.t == 3
        return np.array([self.ep, self.t, float(a[0])]), 0.0, done, {}


class _ToyPolicy:
    def predict(self, s):
        return torch.tensor([s[1] * 0.1, -1.0])


def test_collect_particles_contract():
    states, actions, lens, nxt = dropin.collect_particles(_ToyEnv(), _ToyPolicy(), 3, 5, [0, 1])
    assert states.shape == (3, 6, 3) and states.dtype == np.float32
    assert actions.shape == (3, 5, 2) and actions.dtype == np.float32
    assert lens.dtype == np.int32 and lens.reshape(-1).tolist() == [5, 3, 5]
    assert nxt.shape == (13, 2)
    # trajectory-major order, rows 1..len of every trajectory, filtered columns
    np.testing.assert_array_equal(nxt[:5, 1], [1, 2, 3, 4, 5])
    np.testing.assert_array_equal(nxt[5:8, 0], [1, 1, 1])
    np.testing.assert_array_equal(nxt[5:8, 1], [1, 2, 3])
    # the state after the last step is stored, the rest of a short trajectory stays zero
    assert states[1, 3, 1] == 3 and np.all(states[1, 4:] == 0)


def test_traj_offsets():
    off, n = mepol_b200.traj_offsets(torch.tensor([[5], [3], [5]]), "cpu")
    assert off.tolist() == [0, 5, 8, 13] and n == 13 and off.dtype == torch.int32


def test_install_and_uninstall_patch_the_module():
    ref = types.ModuleType("fake_ref")
    for name in ("collect_particles_and_compute_knn", "compute_importance_weights", "compute_entropy",
                 "compute_kl", "policy_update", "collect_particles"):
        setattr(ref, name, object())
    keep = ref.compute_entropy
    mepol_b200.install(ref)
    assert ref.compute_entropy is mepol_b200.compute_entropy
    assert ref.collect_particles_and_compute_knn is mepol_b200.collect_particles_and_compute_knn
    mepol_b200.uninstall(ref)
    assert ref.compute_entropy is keep


def test_gaussian_log_p_matches_oracle_on_cpu_tensors():
    """The torch expression used on the device is checked here on CPU tensors against the oracle's
    numpy restatement of policy.py:43-61 (same code path, different device)."""
    from oracle import iw_oracle as O
    from conftest import golden_files, load_golden
    g = load_golden(golden_files("iw_ant_like")[0])
    sd = {k[4:]: torch.tensor(v) for k, v in g.items() if k.startswith("tgt.")}
    hidden = [int(h) for h in g["hidden"]]

    class P(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.log_std = torch.nn.Parameter(torch.zeros(int(g["A"]), dtype=torch.float64))
            layers, prev = [], int(g["F"])
            for h in hidden:
                layers += [torch.nn.Linear(prev, h).double(), torch.nn.ReLU()]
                prev = h
            self.net = torch.nn.Sequential(*layers)
            self.mean = torch.nn.Linear(prev, int(g["A"])).double()
    p = P()
    p.load_state_dict(sd)
    s, a, _ = O.flatten_trajectories(g["states"], g["actions"], g["lens"])
    lp = mepol_b200.gaussian_log_p(p, list(p.parameters()), torch.tensor(s, dtype=torch.float64),
                                   torch.tensor(a, dtype=torch.float64))
    np.testing.assert_allclose(lp.detach().numpy(), g["logp_target"], rtol=1e-12, atol=1e-12)


def test_fast_rng_burn_leaves_the_generator_where_the_reference_calls_would():
    """Q7: the shim advances the global CPU generator exactly like the reference's 2*T randn calls."""
    import torch
    from mepol_b200 import entropy as E

    class B:
        lens = [500, 499, 37, 16, 2]
        action_dim = 8
        pinned = {}

    for lens in ([500, 499, 37, 16, 2], [3, 500], [1]):       # the last two contain calls of < 16 elements
        B.lens = lens
        B.pinned = {}
        torch.manual_seed(99)
        for l in lens:
            for _ in range(2):
                torch.randn((l, 8), dtype=torch.float64)
        want = torch.get_rng_state().clone()
        torch.manual_seed(99)
        E._burn_rng(B)
        assert torch.equal(torch.get_rng_state(), want)
    assert E._burn_fast_ok() in (True, False)


def test_rng_burn_interleaved_with_foreign_draws_and_size_changes():
    """The shim's burn equals the literal randn calls inside any sequence of other generator uses."""
    import torch
    from mepol_b200 import entropy as E

    class B:
        action_dim = 8
        pinned = {}

        def __init__(self, lens):
            self.lens = lens
            self.pinned = {}

    def literal(lens):
        for l in lens:
            for _ in range(2):
                torch.randn((l, 8), dtype=torch.float64)

    a, b = B([500, 500, 499]), B([37, 64])
    script = [("burn", a), ("burn", a), ("burn", a), ("draw", None), ("burn", a), ("burn", b), ("burn", a),
              ("burn", a), ("seed", None), ("burn", a), ("burn", a)]
    states = {}
    for mode in ("shim", "literal"):
        torch.manual_seed(5)
        trace = []
        for op, arg in script:
            if op == "burn":
                E._burn_rng(arg) if mode == "shim" else literal(arg.lens)
            elif op == "draw":
                torch.rand(3)
            else:
                torch.manual_seed(11)
            trace.append(torch.get_rng_state().clone())
        states[mode] = trace
    for i, (x, y) in enumerate(zip(states["shim"], states["literal"])):
        assert torch.equal(x, y), f"generator diverged after script step {i}"


def test_attached_handles_stay_bounded():
    """dropin.search_particles attaches one device handle per epoch; every later lookup hits the newest
    entry, so pruning must happen on attach (stale entries pinned device memory forever)."""
    import torch
    from mepol_b200 import entropy as E

    class FakeHandle:
        k1 = 5
    E._handles.clear()
    keep = []
    for i in range(50):
        t = torch.zeros((4, 5), dtype=torch.int64)
        E.attach_handle(t, FakeHandle())
        if i % 10 == 0:
            keep.append(t)                     # a few stay alive, most die at once
        assert len(E._handles) <= E._CACHE_CAP
    del keep
    t = torch.zeros((4, 5), dtype=torch.int64)
    E.attach_handle(t, FakeHandle())
    assert len(E._handles) <= E._CACHE_CAP
    assert E._handles[(id(t), 5)][0]() is t
    E._handles.clear()
