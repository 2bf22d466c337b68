"""Drop-in wiring for the unmodified reference training loop.

`mepol()` (src/algorithms/mepol.py:284-545) looks its helpers up by global name at call time
(:348, :358, :414, :429, :435, :481, :513, :517), so replacing the module attributes before calling
`mepol()` swaps the hot path with zero edits to the reference:

    import src.algorithms.mepol as ref
    import mepol_b200
    mepol_b200.install(ref)          # ref.compute_entropy, ref.compute_kl, ... now run on the B200
    ref.mepol(...)                   # the reference's own training loop

Rollout collection stays on the host CPU (north_star); only lines :190-192 of
collect_particles_and_compute_knn change.
"""
import numpy as np
import torch

from . import entropy as _entropy
from .knn import knn_device, to_device_f32, to_host

_ORIGINALS = {}
_REPLACED = ("collect_particles_and_compute_knn", "compute_importance_weights", "compute_entropy",
             "compute_kl", "policy_update")


def collect_particles(env, policy, num_traj, traj_len, state_filter):
    """Host rollouts with the reference's output contract (mepol.py:70-111): states
    f32[T, L+1, F], actions f32[T, L, A], real_traj_lengths i32[T, 1], next_states f32[N, ns]
    (trajectory-major, rows 1..len of every trajectory, optionally column-filtered)."""
    n_feat = env.num_features
    act_dim = env.action_space.shape[0]
    states = np.zeros((num_traj, traj_len + 1, n_feat), dtype=np.float32)
    actions = np.zeros((num_traj, traj_len, act_dim), dtype=np.float32)
    lengths = np.zeros((num_traj, 1), dtype=np.int32)
    chunks = []
    for ti in range(num_traj):
        obs = env.reset()
        steps = 0
        for t in range(traj_len):
            states[ti, t] = obs
            act = policy.predict(obs).numpy()
            actions[ti, t] = act
            obs, _, done, _ = env.step(act)
            steps = t + 1
            if done:
                break
        states[ti, steps] = obs
        lengths[ti] = steps
        chunks.append(states[ti, 1:steps + 1].reshape(-1, n_feat))
    next_states = np.concatenate(chunks, axis=0)
    if state_filter is not None:
        next_states = next_states[:, state_filter]
    return states, actions, lengths, next_states


def _collector():
    return _ORIGINALS.get("collect_particles", collect_particles)


def search_particles(next_states, k):
    """The replacement for mepol.py:190-192 on a host float32 array.  Returns CPU tensors
    (distances float64 [N,k+1], indices int64 [N,k+1]) plus the device handle that keeps the result
    resident for the objective kernels."""
    x = np.ascontiguousarray(next_states, dtype=np.float32)
    if not np.isfinite(x).all():
        raise ValueError("Input contains NaN or infinity")
    dev = _entropy._device()
    with torch.cuda.device(dev):
        xt = to_device_f32(x, dev)
        res = knn_device(xt, k + 1)
        distances, indices = to_host(res)
    handle = _entropy.SearchHandle(res.indices, res.rk, k + 1)
    _entropy.attach_handle(indices, handle)
    return distances, indices, handle


def collect_particles_and_compute_knn(env, behavioral_policy, num_traj, traj_len,
                                      state_filter, k, num_workers):
    """mepol.py:177-202 with the sklearn fit/kneighbors (:190-192) replaced by the CUDA search."""
    assert num_traj % num_workers == 0, "Please provide a number of trajectories " \
                                        "that can be equally split among workers"
    collect = _collector()
    per_worker = int(num_traj / num_workers)
    if num_workers == 1:
        res = [collect(env, behavioral_policy, per_worker, traj_len, state_filter)]
    else:
        from joblib import Parallel, delayed
        res = Parallel(n_jobs=num_workers)(
            delayed(collect)(env, behavioral_policy, per_worker, traj_len, state_filter)
            for _ in range(num_workers))
    states, actions, real_traj_lengths, next_states = [np.vstack(x) for x in zip(*res)]

    distances, indices, _ = search_particles(next_states, k)

    f64, i64 = torch.float64, torch.int64
    states = torch.tensor(states, dtype=f64)
    actions = torch.tensor(actions, dtype=f64)
    next_states = torch.tensor(next_states, dtype=f64)
    real_traj_lengths = torch.tensor(real_traj_lengths, dtype=i64)
    return states, actions, real_traj_lengths, next_states, distances, indices


def install(ref_module):
    """Monkey-patch the reference's `src.algorithms.mepol` module in place."""
    if not _ORIGINALS:
        for name in _REPLACED + ("collect_particles",):
            if hasattr(ref_module, name):
                _ORIGINALS[name] = getattr(ref_module, name)
    ref_module.collect_particles_and_compute_knn = collect_particles_and_compute_knn
    ref_module.compute_importance_weights = _entropy.compute_importance_weights
    ref_module.compute_entropy = _entropy.compute_entropy
    ref_module.compute_kl = _entropy.compute_kl
    ref_module.policy_update = _entropy.policy_update
    return ref_module


def uninstall(ref_module):
    for name in _REPLACED:
        if name in _ORIGINALS:
            setattr(ref_module, name, _ORIGINALS[name])
    _ORIGINALS.clear()
