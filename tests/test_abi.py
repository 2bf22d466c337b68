"""CPU-side checks of the boundary: the shared library builds/loads, exports every symbol that
include/mepol_b200.h declares, validates arguments, and refuses to compute without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import mepol_b200
from mepol_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "mepol_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mepol_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _header_symbols()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mepol_b200.h but not exported"
    assert set(names) == set(_lib.SIGNATURES), "ctypes table and header disagree"
    assert lib.mepol_version() >= 100


def test_workspace_queries_and_argument_validation():
    assert _lib.knn_workspace_bytes(1000, 1000, 29, 5, _lib.KNN_SIMT) > 0
    assert _lib.rev_csr_workspace_bytes(1000, 4) > 0
    assert _lib.iw_workspace_bytes(1000, 4) > 0
    with pytest.raises(ValueError, match="n_neighbors <= n_samples_fit"):
        _lib.knn_workspace_bytes(3, 3, 29, 5, _lib.KNN_SIMT)
    with pytest.raises(ValueError):
        _lib.knn_workspace_bytes(100, 100, 65, 5, _lib.KNN_SIMT)   # d > MEPOL_MAX_DIM
    with pytest.raises(ValueError):
        _lib.knn_workspace_bytes(100, 100, 8, 61, _lib.KNN_SIMT)   # k1 > MEPOL_MAX_K1
    with pytest.raises(ValueError):
        _lib.knn_workspace_bytes(100, 100, 8, 5, 99)               # unknown algo


def test_null_pointers_are_rejected_without_touching_the_gpu():
    lib = _lib.load()
    rc = lib.mepol_knn_search_f32(None, 10, None, 10, 3, 2, 1, None, None, None, None, 0, None, None)
    assert rc == -1 and b"null" in lib.mepol_last_error()
    rc = lib.mepol_rev_csr(None, 10, 5, None, None, None, 0, None)
    assert rc == -1
    rc = lib.mepol_iw_entropy_fwd_f32(None, None, None, 1, 10, None, 5, None, 3, 0.0, 0.0, 0.0, None, None,
                                      None, None, 0, None)
    assert rc == -1


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_product_path_fails_loudly_without_cuda():
    x = np.zeros((16, 3), dtype=np.float32)
    with pytest.raises(_lib.MepolError, match="no CPU fallback"):
        mepol_b200.knn(x, 4)
    with pytest.raises(_lib.MepolError):
        mepol_b200.knn_device(torch.zeros(16, 3), 5)
    with pytest.raises(_lib.MepolError):
        mepol_b200.compute_entropy(None, None, torch.zeros(1, 2, 3), torch.zeros(1, 1, 1), 1,
                                   torch.ones(1, 1), torch.zeros(1, 5), torch.zeros(1, 5), 4, 1.0, 0.0, 3, 0.0)
    # the C host entry point reports a CUDA error instead of computing on the CPU
    with pytest.raises(_lib.MepolError):
        mepol_b200.knn_host_c(x, 4)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "mepol_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
                assert "import sklearn" not in src and "from sklearn" not in src, f
