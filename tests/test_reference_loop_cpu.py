"""CPU sanity of the reference-loop harness (tests/ref_loop_runner.py): the UNMODIFIED reference `mepol()`
(staged under oracle/_ref by oracle/make_ref.py) runs on the numpy toy environment and writes the three
logs the GPU parity test (tests/test_gpu_reference_loop.py) compares."""
import csv
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_unpatched_reference_loop_runs_on_the_toy_environment(tmp_path):
    sys.path.insert(0, ROOT)
    from oracle import make_ref
    if not make_ref.make(verbose=False):
        pytest.skip("oracle/_ref is not staged (needs /root/reference once)")
    out = tmp_path / "ref"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "ref_loop_runner.py"), "--mode", "ref", "--env",
                        "toy2_ragged", "--out", str(out), "--kl-threshold", "0.02", "--max-off-iters", "4"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "RUN_OK ref toy2_ragged" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    with open(out / "toy2_ragged.csv") as f:
        rows = list(csv.DictReader(f))
    assert [row["epoch"] for row in rows] == ["0", "1", "2"]
    assert all(row["full_entropy"] not in ("", "None", "nan") for row in rows)
    with open(out / "toy2_ragged_off_policy_iter.csv") as f:
        assert len(list(csv.DictReader(f))) >= 1
    for name in ("0-policy", "1-policy", "2-policy", "final_policy.pt", "rng_state.pt"):
        assert (out / name).exists()
    # the staged sources are byte-identical to the reference (SHA256SUMS written at staging time)
    sums = open(os.path.join(ROOT, "oracle", "_ref", "SHA256SUMS")).read().split()
    assert len(sums) == 12
    if os.path.isdir("/root/reference"):
        import hashlib
        for digest, rel in zip(sums[0::2], sums[1::2]):
            with open(os.path.join("/root/reference", rel), "rb") as fh:
                assert hashlib.sha256(fh.read()).hexdigest() == digest, rel
