"""Stage the reference's own hot-path sources under oracle/_ref/ (TEST INFRASTRUCTURE).

    python oracle/make_ref.py        # needs /root/reference; a no-op message otherwise

The reference is pure Python: "building" it means copying the three files the hot path lives in
(byte for byte, nothing is edited) next to their package markers:

    /root/reference/src/algorithms/mepol.py   -> oracle/_ref/src/algorithms/mepol.py
    /root/reference/src/policy.py             -> oracle/_ref/src/policy.py
    /root/reference/src/utils/dtypes.py       -> oracle/_ref/src/utils/dtypes.py

oracle/_ref/ is git-ignored (no reference source enters the history) but not gpurun-ignored, so the
unmodified reference travels to the GPU box, where /root/reference does not exist.  It is used
  * by tests/test_gpu_reference_loop.py: the real `mepol()` training loop (mepol.py:284-545) run
    un-patched and through `mepol_b200.install()` on the same seeds, and
  * by bench.py's CPU arm (`cpu_baseline.kind == "reference"`): the reference's own
    compute_entropy / backward / compute_kl timed on the box's host cores.
Only tests/, __graft_entry__ and bench.py's CPU legs may import it; the product never does.
"""
import hashlib
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference"
DST = os.path.join(HERE, "_ref")
FILES = ["src/__init__.py", "src/algorithms/__init__.py", "src/algorithms/mepol.py", "src/policy.py",
         "src/utils/__init__.py", "src/utils/dtypes.py"]


def available():
    return all(os.path.exists(os.path.join(DST, f)) for f in FILES)


def make(verbose=True):
    """Copy the files (idempotent).  Returns True when oracle/_ref is usable afterwards."""
    if not os.path.isdir(REF_SRC):
        if verbose:
            print(f"make_ref: {REF_SRC} is absent; keeping the staged copy (available={available()})")
        return available()
    lines = []
    for f in FILES:
        src, dst = os.path.join(REF_SRC, f), os.path.join(DST, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as fh:
            lines.append(f"{hashlib.sha256(fh.read()).hexdigest()}  {f}")
    with open(os.path.join(DST, "SHA256SUMS"), "w") as fh:
        fh.write("\n".join(lines) + "\n")
    if verbose:
        print(f"make_ref: staged {len(FILES)} files under {DST}")
    return True


def import_reference():
    """(src.algorithms.mepol, src.policy) of the staged, unmodified reference.  matplotlib is absent
    from the image and only get_heatmap (mepol.py:19-67, out of scope) uses it: it is stubbed.
    NB: importing src.utils.dtypes switches torch's default dtype to float64 process-wide
    (dtypes.py:3) -- callers run this in a subprocess or restore the default themselves."""
    if not available():
        raise RuntimeError("oracle/_ref is not staged: run `python oracle/make_ref.py` where /root/reference exists")
    for name in ("matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                m = types.ModuleType(name)
                m.use = lambda *a, **k: None
                sys.modules[name] = m
    if "matplotlib.pyplot" in sys.modules and not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if DST not in sys.path:
        sys.path.insert(0, DST)
    import warnings
    warnings.filterwarnings("ignore")
    import src.algorithms.mepol as ref_mepol
    import src.policy as ref_policy
    return ref_mepol, ref_policy


if __name__ == "__main__":
    ok = make()
    sys.exit(0 if ok else 1)
