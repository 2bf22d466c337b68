"""Bring-up of knn_tc2.cu: the raw fp16 accumulator tile against a float64 emulation of the packed
fp16 operands.  Reports the error in units of the fp16 ulp of the result (rounding behaviour of the
tensor core's fp16 accumulation) and against the bound the kernel's band assumes."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mepol_b200 import _lib


def operands(x):
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
    return xb, h, m, l, sigma


def ulp16(v):
    a = np.maximum(np.abs(v), 2.0 ** -14)
    return 2.0 ** (np.floor(np.log2(a)) - 10)


def run(n, d, seed, kind="normal"):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d)).astype(np.float32)
    if kind == "traj":
        x = np.cumsum(0.05 * x, axis=0).astype(np.float32)
    if kind == "offset":
        x = (x + 100.0).astype(np.float32)
    xt = torch.from_numpy(x).cuda()
    out = torch.full((256, 256), float("nan"), dtype=torch.float32, device="cuda")
    ws = torch.empty(_lib.knn_workspace_bytes(n, n, d, 5, _lib.KNN_TC), dtype=torch.uint8, device="cuda")
    rc = _lib.load().mepol_debug_tc2_scores(xt.data_ptr(), n, xt.data_ptr(), n, d, 5, out.data_ptr(), ws.data_ptr(),
                                            ws.numel(), torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "mepol_debug_tc2_scores")
    torch.cuda.synchronize()
    got = out.cpu().numpy().astype(np.float64)
    xb, h, m, l, sigma = operands(x)
    q, r = xb[:256], xb[:256]
    k_first = min(d, 16)
    d1 = (-2.0 * (q[:, :k_first] @ r[:, :k_first].T)).numpy()
    if d + 3 <= 16:
        exact = ((h + m + l)[:256][None, :] - 2.0 * (q @ r.T)).numpy()
        d1 = np.zeros_like(exact)
    else:
        nrm = [h, m, l]
        first_norm = sum(nrm[i][:256] for i in range(3) if d + i < 16) if d < 16 else 0.0
        d1 = d1 + (first_norm[None, :].numpy() if d < 16 else 0.0)
        exact = ((h + m + l)[:256][None, :] - 2.0 * (q @ r.T)).numpy()
    err = np.abs(got - exact)
    # two roundings: ulp of the intermediate + ulp of the result
    bound = ulp16(d1) * (d + 3 > 16) + ulp16(exact)
    rel = err / bound
    print(f"n={n} d={d} {kind}: sigma={sigma:g} nan={np.isnan(got).sum()} max|err|={err.max():.4g} "
          f"max err/(ulp(D1)+ulp(D))={rel.max():.3f} mean={rel.mean():.3f} max|D|={np.abs(exact).max():.1f} "
          f"max|D1|={np.abs(d1).max():.1f}")
    # same-value emulation with round-to-nearest at both steps
    e1 = torch.from_numpy(d1).to(torch.float16).double().numpy() if d + 3 > 16 else d1
    rn = torch.from_numpy(e1 + (exact - d1)).to(torch.float16).double().numpy()
    print(f"   exact match with RN(RN(D1)+rest): {np.mean(rn == got) * 100:.2f}%   max|got-rn|/ulp={np.max(np.abs(got - rn) / ulp16(exact)):.3f}")
    if os.environ.get("EXP_VERBOSE"):
        bad = np.argwhere(rn != got)
        order = np.argsort(-rel[bad[:, 0], bad[:, 1]])[:6]
        for i, j in bad[order]:
            print(f"      [{i},{j}] D1={d1[i, j]:.6f} RN16(D1)={e1[i, j]:.6f} rest={exact[i, j] - d1[i, j]:.6f} exact={exact[i, j]:.6f} "
                  f"got={got[i, j]:.6f} rn={rn[i, j]:.6f} err/bound={rel[i, j]:.3f}")
    return rel.max()


if __name__ == "__main__":
    worst = 0.0
    cases = [(1024, 29, "normal"), (4096, 29, "traj"), (2048, 24, "normal"), (2048, 16, "normal"),
             (2048, 14, "normal"), (2048, 13, "normal"), (1024, 7, "normal"), (2048, 29, "offset")]
    if os.environ.get("EXP_VERBOSE"):
        cases = [(2048, d, "normal") for d in range(14, 30)] + [(3000 + 7 * s, 29, "normal") for s in range(6)]
    for (n, d, kind) in cases:
        worst = max(worst, run(n, d, n + d, kind))
    print("WORST err/(ulp(D1)+ulp(D)) =", worst, "-> the band assumes <= 1.0")
