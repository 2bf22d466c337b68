"""Build recipe of libmepol_b200.so (hand-written sm_100a CUDA + the C ABI of include/mepol_b200.h).

    python -m mepol_b200.build [--force]

Plain nvcc, in-tree output (mepol_b200/libmepol_b200.so travels to the GPU box with the snapshot).
nvcc cross-compiles sm_100a without a GPU, so this runs in the CPU-only build container.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libmepol_b200.so")
SOURCES = ["capi.cu", "knn_prep.cu", "knn_simt.cu", "knn_tc.cu", "knn_tc2.cu", "knn_grid.cu", "knn_rerank.cu", "iw_entropy.cu", "gauss_logp.cu"]
NVCC = os.environ.get("MEPOL_NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr",
         "-ccbin", "/usr/bin/g++"] + os.environ.get("MEPOL_NVCC_EXTRA", "").split()


def _deps():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(HERE, "..", "include", "mepol_b200.h"))
    return max(os.path.getmtime(h) for h in hdrs)


def _compile(src, force, verbose):
    s = os.path.join(CSRC, src)
    o = os.path.join(OBJ, src.replace(".cu", ".o"))
    if not force and os.path.exists(o) and os.path.getmtime(o) >= max(os.path.getmtime(s), _deps()):
        return o, False, ""
    r = subprocess.run([NVCC, *FLAGS, "-c", s, "-o", o], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    log = r.stdout + r.stderr
    with open(o + ".ptxas.log", "w") as f:
        f.write(log)
    return o, True, log


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        res = list(ex.map(lambda s: _compile(s, force, verbose), SOURCES))
    objs = [r[0] for r in res]
    if verbose:
        for _, did, log in res:
            if did:
                sys.stdout.write(log)
    if force or any(r[1] for r in res) or not os.path.exists(LIB):
        r = subprocess.run([NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
                            "-ccbin", "/usr/bin/g++", "-lcudart"],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
