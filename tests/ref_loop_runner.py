"""Runs the reference's OWN training loop `mepol()` (oracle/_ref/src/algorithms/mepol.py:284-545, staged
unmodified by oracle/make_ref.py) on a deterministic numpy toy environment, either un-patched
(`--mode ref`: sklearn + torch CPU float64, the reference as shipped) or after
`mepol_b200.install(ref)` (`--mode ours`: the same loop, the hot path on the B200).

TEST INFRASTRUCTURE (used by tests/test_gpu_reference_loop.py in a subprocess, because importing the
reference switches torch's default dtype process-wide, dtypes.py:3, and `mepol()` may call exit()).

Writes into --out: the reference's own CSV logs ({env}.csv, {env}_off_policy_iter.csv), the saved
policies, and final_policy.pt (state_dict returned by mepol()).
"""
import argparse
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


class ToyEnv:
    """Smooth stochastic point dynamics with the interface `mepol()` needs (num_features,
    action_space.shape, seed, reset, step): s' = 0.98 s + 0.1 tanh(W a) + 0.01 N(0, I).
    Continuous noise keeps the particles tie-free.  `ragged` ends episodes early at seeded random
    lengths (real_traj_lengths then differ per trajectory, mepol.py:92-96)."""

    def __init__(self, num_features, action_dim, ragged=False):
        self.num_features = num_features
        self.action_space = types.SimpleNamespace(shape=(action_dim,))
        self.W = np.random.RandomState(1234).standard_normal((num_features, action_dim)) / np.sqrt(action_dim)
        self.ragged = ragged
        self.rng = np.random.RandomState(0)
        self.s = np.zeros(num_features)
        self.t = 0
        self.horizon = 1 << 30

    def seed(self, seed):
        self.rng = np.random.RandomState(seed)

    def reset(self):
        self.s = 0.05 * self.rng.standard_normal(self.num_features)
        self.t = 0
        self.horizon = int(self.rng.randint(24, 64)) if self.ragged else 1 << 30
        return self.s.copy()

    def step(self, a):
        a = np.asarray(a, dtype=np.float64).reshape(-1)
        self.s = 0.98 * self.s + 0.1 * np.tanh(self.W @ a) + 0.01 * self.rng.standard_normal(self.num_features)
        self.t += 1
        return self.s.copy(), 0.0, self.t >= self.horizon, {}


ENVS = {
    # name: (F, A, state_filter, hidden, ragged, k)
    "toy2": (2, 2, None, [32, 32], False, 4),
    "toy29": (29, 8, list(range(7)), [64, 48], False, 4),
    "toy2_ragged": (2, 1, None, [32], True, 3),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", required=True, choices=["ref", "ours"])
    ap.add_argument("--env", required=True, choices=sorted(ENVS))
    ap.add_argument("--out", required=True)
    ap.add_argument("--epochs", type=int, default=2)
    ap.add_argument("--num-traj", type=int, default=8)
    ap.add_argument("--traj-len", type=int, default=64)
    ap.add_argument("--max-off-iters", type=int, default=4)
    ap.add_argument("--kl-threshold", type=float, default=0.02)
    ap.add_argument("--lr", type=float, default=2e-3)
    ap.add_argument("--seed", type=int, default=7)
    ap.add_argument("--compute-dtype", default="float64", choices=["float64", "float32"])
    args = ap.parse_args()

    import torch
    from oracle import make_ref
    ref, ref_policy = make_ref.import_reference()
    F, A, state_filter, hidden, ragged, k = ENVS[args.env]
    env = ToyEnv(F, A, ragged)
    os.makedirs(args.out, exist_ok=True)

    if args.mode == "ours":
        import mepol_b200
        mepol_b200.configure(compute_dtype=getattr(torch, args.compute_dtype))
        mepol_b200.install(ref)

    def create_policy(is_behavioral=False):
        return ref_policy.GaussianPolicy(num_features=F, hidden_sizes=hidden, action_dim=A)

    final = ref.mepol(
        env=env, env_name=args.env, state_filter=state_filter, create_policy=create_policy, k=k,
        kl_threshold=args.kl_threshold, max_off_iters=args.max_off_iters, use_backtracking=True,
        backtrack_coeff=2, max_backtrack_try=3, eps=0.0, learning_rate=args.lr, num_traj=args.num_traj,
        traj_len=args.traj_len, num_epochs=args.epochs, optimizer="adam", full_entropy_traj_scale=2,
        full_entropy_k=k, heatmap_every=1, heatmap_discretizer=None, heatmap_episodes=0, heatmap_num_steps=0,
        heatmap_cmap=None, heatmap_labels=None, heatmap_interp=None, seed=args.seed, out_path=args.out,
        num_workers=1)
    torch.save(final.state_dict(), os.path.join(args.out, "final_policy.pt"))
    # the generator state after the run: equal states mean the drop-in consumed exactly the
    # reference's random draws, so a longer run would stay aligned too
    torch.save(torch.get_rng_state(), os.path.join(args.out, "rng_state.pt"))
    if args.mode == "ours":
        from mepol_b200 import _lib
        print("native library:", _lib.LIB_PATH)
    print("RUN_OK", args.mode, args.env)


if __name__ == "__main__":
    main()
