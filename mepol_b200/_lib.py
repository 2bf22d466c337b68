"""ctypes binding of libmepol_b200.so (the C ABI declared in include/mepol_b200.h).

The product path has no CPU fallback: if the library is missing or an entry point fails, the error
is raised -- never swallowed."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmepol_b200.so")

KNN_AUTO, KNN_SIMT, KNN_TC, KNN_GRID = 0, 1, 2, 3
KNN_NSTATS = 8
MAX_DIM, MAX_K1 = 64, 60

_c = ctypes
_vp, _i64, _i32, _dbl, _sz = _c.c_void_p, _c.c_int64, _c.c_int, _c.c_double, _c.c_size_t

# name -> (restype, argtypes); mirrors include/mepol_b200.h one to one
SIGNATURES = {
    "mepol_last_error": (_c.c_char_p, []),
    "mepol_version": (_i32, []),
    "mepol_device_info": (_i32, [_c.POINTER(_i32)] * 3),
    "mepol_profile_enable": (_i32, [_i32]),
    "mepol_profile_read": (_i32, [_c.POINTER(_c.c_float), _i32, _c.POINTER(_i32)]),
    "mepol_knn_workspace_bytes": (_i32, [_i64, _i64, _i32, _i32, _i32, _c.POINTER(_sz)]),
    "mepol_knn_search_f32": (_i32, [_vp, _i64, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _sz, _vp, _vp]),
    "mepol_debug_tc_scores": (_i32, [_vp, _i64, _vp, _i64, _i32, _i32, _vp, _vp, _sz, _vp]),
    "mepol_debug_tc2_scores": (_i32, [_vp, _i64, _vp, _i64, _i32, _i32, _vp, _vp, _sz, _vp]),
    "mepol_knn_search_host_f32": (_i32, [_vp, _i64, _i32, _i32, _i32, _vp, _vp]),
    "mepol_rev_csr_workspace_bytes": (_i32, [_i64, _i32, _c.POINTER(_sz)]),
    "mepol_rev_csr": (_i32, [_vp, _i64, _i32, _vp, _vp, _vp, _sz, _vp]),
    "mepol_iw_workspace_bytes": (_i32, [_i64, _i64, _c.POINTER(_sz)]),
    "mepol_iw_entropy_fwd_f32": (_i32, [_vp, _vp, _vp, _i64, _i64, _vp, _i32, _vp, _i32, _dbl, _dbl, _dbl,
                                        _vp, _vp, _vp, _vp, _sz, _vp]),
    "mepol_iw_entropy_fwd_f64": (_i32, [_vp, _vp, _vp, _i64, _i64, _vp, _i32, _vp, _i32, _dbl, _dbl, _dbl,
                                        _vp, _vp, _vp, _vp, _sz, _vp]),
    "mepol_iw_entropy_fwd_bwd_f32": (_i32, [_vp, _vp, _vp, _i64, _i64, _vp, _i32, _vp, _i32, _dbl, _dbl, _dbl,
                                            _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "mepol_iw_entropy_fwd_bwd_f64": (_i32, [_vp, _vp, _vp, _i64, _i64, _vp, _i32, _vp, _i32, _dbl, _dbl, _dbl,
                                            _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "mepol_iw_entropy_bwd_f32": (_i32, [_vp, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp]),
    "mepol_iw_entropy_bwd_f64": (_i32, [_vp, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp]),
    "mepol_gauss_logp_workspace_bytes": (_i32, [_i64, _i32, _c.POINTER(_sz)]),
    "mepol_gauss_logp_fwd_f32": (_i32, [_vp, _vp, _vp, _i64, _i32, _vp, _vp]),
    "mepol_gauss_logp_fwd_f64": (_i32, [_vp, _vp, _vp, _i64, _i32, _vp, _vp]),
    "mepol_gauss_logp_bwd_f32": (_i32, [_vp, _vp, _vp, _vp, _i64, _i32, _vp, _vp, _vp, _sz, _vp]),
    "mepol_gauss_logp_bwd_f64": (_i32, [_vp, _vp, _vp, _vp, _i64, _i32, _vp, _vp, _vp, _sz, _vp]),
}

_lib = None


class MepolError(RuntimeError):
    pass


def load():
    """Load the shared library (raises if it has not been built: python -m mepol_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MepolError(f"{LIB_PATH} is missing -- build it with `python -m mepol_b200.build` "
                             "(there is no CPU fallback)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc, what):
    if rc != 0:
        msg = load().mepol_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise ValueError(f"{what}: {msg}")
        raise MepolError(f"{what} failed (rc={rc}): {msg}")


def knn_workspace_bytes(n_q, n_ref, d, k1, algo=KNN_AUTO):
    out = _sz(0)
    check(load().mepol_knn_workspace_bytes(n_q, n_ref, d, k1, algo, ctypes.byref(out)), "mepol_knn_workspace_bytes")
    return out.value


def rev_csr_workspace_bytes(n, k):
    out = _sz(0)
    check(load().mepol_rev_csr_workspace_bytes(n, k, ctypes.byref(out)), "mepol_rev_csr_workspace_bytes")
    return out.value


def iw_workspace_bytes(n, n_traj):
    out = _sz(0)
    check(load().mepol_iw_workspace_bytes(n, n_traj, ctypes.byref(out)), "mepol_iw_workspace_bytes")
    return out.value


def gauss_logp_workspace_bytes(n, A):
    out = _sz(0)
    check(load().mepol_gauss_logp_workspace_bytes(n, A, ctypes.byref(out)), "mepol_gauss_logp_workspace_bytes")
    return out.value


def profile_enable(on=True):
    check(load().mepol_profile_enable(1 if on else 0), "mepol_profile_enable")


def profile_read(max_n=64):
    """Per-launch milliseconds of the pruning kernel for the searches issued since profile_enable."""
    buf = (_c.c_float * max_n)()
    n = _i32(0)
    check(load().mepol_profile_read(buf, max_n, ctypes.byref(n)), "mepol_profile_read")
    return [buf[i] for i in range(n.value)]


def prune_kernel_name(algo, d, k1, n_ref):
    """Name of the kernel mepol_knn_search_f32 spends its time in for this problem (mirrors the AUTO rule
    of csrc/capi.cu::knn_make_plan and the tensor-core variant switch) -- for bench.py's roofline record."""
    if algo == KNN_AUTO:
        if d <= 3:
            algo = KNN_GRID
        else:
            algo = KNN_TC if (d + 3 <= 32 and 2 * k1 + 8 <= 32 and n_ref >= 512 and d >= 4) else KNN_SIMT
    if algo == KNN_GRID:
        return "knn_grid_search_kernel"
    if algo == KNN_SIMT:
        return "knn_prune_simt_kernel"
    import os
    v = os.environ.get("MEPOL_TC_VARIANT", "")
    return "knn_prune_tc_kernel" if v in ("0", "1", "2") else "knn_prune_tc2_kernel"
