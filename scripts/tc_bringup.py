"""Bring-up diagnostics for the tcgen05 pruner (run on the B200 box; the MEPOL_TC_EPI_MODE timing
switch needs a library built with -DMEPOL_BRINGUP: MEPOL_NVCC_EXTRA=-DMEPOL_BRINGUP python -m mepol_b200.build --force):
   python scripts/tc_bringup.py dump      # raw accumulator tile vs a numpy emulation of the fp16 operands
   python scripts/tc_bringup.py knn N     # TC search vs SIMT search (indices must be identical), timing
"""
import ctypes
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mepol_b200
from mepol_b200 import _lib


def expected_scores(x, nq=256, nr=128):
    xd = torch.from_numpy(x).double()
    mean = (xd.sum(0) / x.shape[0]).float().double()
    ext = torch.maximum(xd.max(0).values - mean, mean - xd.min(0).values)
    sigma = 2.0 ** np.floor(np.log2(128.0 / float(ext.pow(2).sum().sqrt())))
    xc = (xd - mean) * sigma
    xb = xc.to(torch.float16).double()
    n2 = (xc * xc).sum(1)
    h = n2.to(torch.float16).double()
    m = (n2 - h).to(torch.float16).double()
    l = (n2 - h - m).to(torch.float16).double()
    nsum = h + m + l
    q, r = xb[:nq], xb[:nr]
    return (nsum[:nr][None, :] - 2.0 * (q @ r.T)).numpy(), n2.numpy()


def dump():
    rng = np.random.default_rng(0)
    n, d = 1024, 29
    x = rng.standard_normal((n, d)).astype(np.float32)
    xt = torch.from_numpy(x).cuda()
    out = torch.full((256, 128), float("nan"), dtype=torch.float32, device="cuda")
    need = _lib.knn_workspace_bytes(n, n, d, 5, _lib.KNN_TC)
    ws = torch.empty(need, dtype=torch.uint8, device="cuda")
    rc = _lib.load().mepol_debug_tc_scores(xt.data_ptr(), n, xt.data_ptr(), n, d, 5, out.data_ptr(), ws.data_ptr(),
                                           ws.numel(), torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "mepol_debug_tc_scores")
    torch.cuda.synchronize()
    got = out.cpu().numpy().astype(np.float64)
    exp, n2 = expected_scores(x)
    err = np.abs(got - exp)
    print("swap=%s  max|err|=%.4g  mean|err|=%.4g  nan=%d" % (os.environ.get("MEPOL_TC_SWAP_LBO_SBO", "0"),
                                                               np.nanmax(err), np.nanmean(err), np.isnan(got).sum()))
    print("got[0,:6] =", got[0, :6])
    print("exp[0,:6] =", exp[0, :6])
    print("got[1,:6] =", got[1, :6])
    print("exp[1,:6] =", exp[1, :6])
    print("got[130,:4] =", got[130, :4], " exp:", exp[130, :4])
    ok = np.nanmax(err) < 2e-3 and not np.isnan(got).any()
    print("DUMP", "OK" if ok else "MISMATCH")
    return ok


def knn(n):
    rng = np.random.default_rng(1)
    x = rng.standard_normal((n, 29)).astype(np.float32)
    xt = torch.from_numpy(x).cuda()
    a = mepol_b200.knn_device(xt, 5, algo=_lib.KNN_SIMT)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    b = mepol_b200.knn_device(xt, 5, algo=_lib.KNN_TC)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    _lib.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    b = mepol_b200.knn_device(xt, 5, algo=_lib.KNN_TC)
    e1.record()
    torch.cuda.synchronize()
    prune = _lib.profile_read()
    same = torch.equal(a.indices, b.indices) and torch.equal(a.distances, b.distances)
    print("N=%d  identical=%s  tc_search_ms=%.2f (first %.1f)  prune_ms=%s  stats=%s simt_stats=%s" % (
        n, same, e0.elapsed_time(e1), (t1 - t0) * 1e3, prune, b.stats_dict(), a.stats_dict()))
    if not same:
        bad = (a.indices != b.indices).any(1).nonzero().flatten()
        print("rows differing:", bad.numel(), bad[:10].tolist())
        r = int(bad[0])
        print(a.indices[r].tolist(), b.indices[r].tolist(), a.distances[r].tolist(), b.distances[r].tolist())
    return same


def time_only(n):
    rng = np.random.default_rng(1)
    x = rng.standard_normal((n, 29)).astype(np.float32)
    xt = torch.from_numpy(x).cuda()
    mepol_b200.knn_device(xt, 5, algo=_lib.KNN_TC)
    _lib.profile_enable(True)
    for _ in range(3):
        mepol_b200.knn_device(xt, 5, algo=_lib.KNN_TC)
    torch.cuda.synchronize()
    print("mode=%s N=%d prune_ms=%s" % (os.environ.get("MEPOL_TC_EPI_MODE", "0"), n, _lib.profile_read()))
    return True


if __name__ == "__main__":
    if sys.argv[1] == "time":
        sys.exit(0 if time_only(int(sys.argv[2])) else 1)
    if sys.argv[1] == "dump":
        sys.exit(0 if dump() else 1)
    else:
        sys.exit(0 if knn(int(sys.argv[2])) else 1)
