"""mepol_b200 -- B200-native (sm_100a) implementation of MEPOL's k-NN state-entropy hot path.

Public surface (mirrors /root/reference/src/algorithms/mepol.py):
    knn(next_states, k)                       -> (distances, indices)      replaces :190-192
    collect_particles_and_compute_knn(...)    same signature as :177-202
    compute_importance_weights / compute_entropy / compute_kl / policy_update   same as :114-174, :268-281
    install(ref_module) / uninstall(ref_module)   monkey-patch the reference in place
All compute runs in libmepol_b200.so (hand-written CUDA behind the C ABI of include/mepol_b200.h);
there is no CPU fallback.
"""
from . import _lib
from ._lib import KNN_AUTO, KNN_GRID, KNN_SIMT, KNN_TC, MepolError
from .knn import KnnResult, knn, knn_device, knn_host_c
from .entropy import (IWEntropyFunction, SearchHandle, compute_entropy, compute_importance_weights,
                      compute_kl, configure, gaussian_log_p, iw_entropy, policy_update, traj_offsets)
from .dropin import (collect_particles, collect_particles_and_compute_knn, install, search_particles,
                     uninstall)

__all__ = [
    "KNN_AUTO", "KNN_GRID", "KNN_SIMT", "KNN_TC", "MepolError", "KnnResult", "knn", "knn_device", "knn_host_c",
    "IWEntropyFunction", "SearchHandle", "compute_entropy", "compute_importance_weights", "compute_kl",
    "configure", "gaussian_log_p", "iw_entropy", "policy_update", "traj_offsets", "collect_particles",
    "collect_particles_and_compute_knn", "install", "search_particles", "uninstall",
]
