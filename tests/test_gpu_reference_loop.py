"""The drop-in under the reference's OWN caller.

The unmodified training loop `mepol()` (reference src/algorithms/mepol.py:284-545; call sites of the hot
path at :348, :358, :414, :429, :435, :481, :513, :517) is run twice with the same seed on a
deterministic numpy toy environment -- un-patched (sklearn + torch CPU float64) and after
`mepol_b200.install(ref)` -- and everything the loop logs or returns is compared:

  * {env}.csv: per-epoch loss, entropy, full-entropy (the 2x batch evaluated with the same policy
    object passed twice under no_grad, :348-354 / :512-519) and the number of accepted off-policy steps;
  * {env}_off_policy_iter.csv: entropy, KL and learning rate of every accepted off-policy step
    (same count => the same accept / backtrack decisions were taken);
  * the final policy's state_dict and the saved per-epoch policies;
  * torch's global generator state after the run (quirk Q7: the drop-in burns the reference's draws).

The reference sources come from oracle/_ref (staged byte for byte by oracle/make_ref.py where
/root/reference exists; git-ignored, travels to the GPU box).  Each run is a subprocess: importing
the reference flips torch's default dtype process-wide (dtypes.py:3).
"""
import csv
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUNNER = os.path.join(ROOT, "tests", "ref_loop_runner.py")

pytestmark = pytest.mark.gpu

# entropy / KL of a step agree to ~1e-13 (float64 on both sides, different summation orders); the
# tolerance leaves room for the amplification through a few Adam steps
RTOL, ATOL = 1e-8, 1e-10


def _staged():
    sys.path.insert(0, ROOT)
    from oracle import make_ref
    return make_ref.make(verbose=False)


def _run(mode, env, out, extra):
    cmd = [sys.executable, RUNNER, "--mode", mode, "--env", env, "--out", str(out), *extra]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0 and "RUN_OK" in r.stdout, f"{mode} run failed:\n{r.stdout[-3000:]}\n{r.stderr[-3000:]}"
    return r.stdout


def _rows(path):
    with open(path) as f:
        return list(csv.DictReader(f))


CASES = [
    # env, runner arguments (kl threshold low enough on toy2 that backtracking is exercised)
    ("toy2", ["--kl-threshold", "0.02", "--max-off-iters", "4"]),
    ("toy29", ["--kl-threshold", "0.2", "--max-off-iters", "4", "--lr", "1e-3"]),
    ("toy2_ragged", ["--kl-threshold", "0.02", "--max-off-iters", "4"]),
]


@pytest.mark.parametrize("env,extra", CASES, ids=[c[0] for c in CASES])
def test_reference_loop_patched_equals_unpatched(env, extra, tmp_path):
    if not _staged():
        pytest.skip("oracle/_ref is not staged (run python oracle/make_ref.py where /root/reference exists)")
    import torch
    out_ref, out_ours = tmp_path / "ref", tmp_path / "ours"
    _run("ref", env, out_ref, extra)
    log = _run("ours", env, out_ours, extra)
    assert "libmepol_b200.so" in log

    ep_ref, ep_ours = _rows(out_ref / f"{env}.csv"), _rows(out_ours / f"{env}.csv")
    assert len(ep_ref) == len(ep_ours) == 3                      # epoch 0 + 2 training epochs
    for a, b in zip(ep_ref, ep_ours):
        assert a["epoch"] == b["epoch"] and a["num_off_iters"] == b["num_off_iters"]
        for key in ("loss", "entropy", "full_entropy"):
            np.testing.assert_allclose(float(b[key]), float(a[key]), rtol=RTOL, atol=ATOL, err_msg=f"{key} epoch {a['epoch']}")

    it_ref, it_ours = _rows(out_ref / f"{env}_off_policy_iter.csv"), _rows(out_ours / f"{env}_off_policy_iter.csv")
    assert len(it_ref) == len(it_ours), "different number of accepted off-policy iterations"
    assert len(it_ref) >= 1
    for a, b in zip(it_ref, it_ours):
        assert (a["epoch"], a["off_policy_iter"]) == (b["epoch"], b["off_policy_iter"])
        assert float(a["learning_rate"]) == float(b["learning_rate"]), "different backtracking decisions"
        for key in ("entropy", "kl"):
            np.testing.assert_allclose(float(b[key]), float(a[key]), rtol=RTOL, atol=ATOL, err_msg=f"{key} at {a}")

    for name in ("final_policy.pt", "0-policy", "1-policy", "2-policy"):
        sd_ref = torch.load(out_ref / name, weights_only=True)
        sd_ours = torch.load(out_ours / name, weights_only=True)
        assert sd_ref.keys() == sd_ours.keys()
        for key in sd_ref:
            np.testing.assert_allclose(sd_ours[key].numpy(), sd_ref[key].numpy(), rtol=1e-7, atol=1e-9,
                                       err_msg=f"{name}:{key}")
    # same random draws consumed => identical rollouts from here on
    assert torch.equal(torch.load(out_ref / "rng_state.pt", weights_only=True),
                       torch.load(out_ours / "rng_state.pt", weights_only=True))


def test_reference_loop_float32_kernels(tmp_path):
    """Same loop with the bandwidth-optimal float32 objective kernels: north_star's 1e-5 bar on the
    logged entropy / KL of the first accepted step (later steps inherit optimiser drift)."""
    if not _staged():
        pytest.skip("oracle/_ref is not staged")
    extra = ["--kl-threshold", "0.2", "--max-off-iters", "3", "--lr", "1e-3", "--epochs", "1"]
    out_ref, out_ours = tmp_path / "ref", tmp_path / "ours"
    _run("ref", "toy29", out_ref, extra)
    _run("ours", "toy29", out_ours, extra + ["--compute-dtype", "float32"])
    a, b = _rows(out_ref / "toy29.csv"), _rows(out_ours / "toy29.csv")
    assert len(a) == len(b) == 2
    np.testing.assert_allclose(float(b[0]["entropy"]), float(a[0]["entropy"]), rtol=1e-5)
    np.testing.assert_allclose(float(b[0]["full_entropy"]), float(a[0]["full_entropy"]), rtol=1e-5)
    ia, ib = _rows(out_ref / "toy29_off_policy_iter.csv"), _rows(out_ours / "toy29_off_policy_iter.csv")
    assert len(ia) >= 1 and len(ib) >= 1
    np.testing.assert_allclose(float(ib[0]["entropy"]), float(ia[0]["entropy"]), rtol=1e-5)
    np.testing.assert_allclose(float(ib[0]["kl"]), float(ia[0]["kl"]), rtol=1e-3, atol=1e-7)
