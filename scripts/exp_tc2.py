"""Bring-up of knn_tc2.cu: the raw fp16 accumulator tile against a float64 emulation of the packed fp16
operands.  Separates the two error sources of the tensor core's fp16 accumulation:
  rounding of the accumulator after each K = 16 step (in ulps of |D1| and |D|), and
  the internal alignment of the 17 addends of a step (relative to the largest addend).
    python scripts/exp_tc2.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mepol_b200 import _lib


def ulp16(v):
    a = np.maximum(np.abs(v), 2.0 ** -14)
    return 2.0 ** (np.floor(np.log2(a)) - 10)


def operands(x):
    xd = x.astype(np.float64)
    mean = (xd.sum(0) / x.shape[0]).astype(np.float32).astype(np.float64)
    ext = np.maximum(xd.max(0) - mean, mean - xd.min(0))
    sigma = 2.0 ** np.floor(np.log2(128.0 / np.sqrt((ext ** 2).sum())))
    xc = (xd - mean) * sigma
    xb = xc.astype(np.float16).astype(np.float64)
    n2 = (xc * xc).sum(1)
    h = n2.astype(np.float16).astype(np.float64)
    m = (n2 - h).astype(np.float16).astype(np.float64)
    l = (n2 - h - m).astype(np.float16).astype(np.float64)
    return xb, [h, m, l], sigma


def run(n, d, seed, kind="normal"):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d))
    if kind == "traj":
        x = np.cumsum(0.05 * x, axis=0)
    if kind == "offset":
        x = x + 100.0
    if kind == "outliers":
        x[:40] *= 6.0
    x = x.astype(np.float32)
    xt = torch.from_numpy(x).cuda()
    out = torch.full((256, 256), float("nan"), dtype=torch.float32, device="cuda")
    ws = torch.empty(_lib.knn_workspace_bytes(n, n, d, 5, _lib.KNN_TC), dtype=torch.uint8, device="cuda")
    rc = _lib.load().mepol_debug_tc2_scores(xt.data_ptr(), n, xt.data_ptr(), n, d, 5, out.data_ptr(), ws.data_ptr(),
                                            ws.numel(), torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "mepol_debug_tc2_scores")
    torch.cuda.synchronize()
    got = out.cpu().numpy().astype(np.float64)
    xb, nrm, sigma = operands(x)
    q, r = xb[:256], xb[:256]
    cq = np.concatenate([q, np.ones((256, 3))], axis=1)
    cr = np.concatenate([-2.0 * r, np.stack([nv[:256] for nv in nrm], axis=1)], axis=1)
    two = d + 3 > 16
    terms = cq[:, None, :] * cr[None, :, :]                    # [256, 256, d+3] exact products
    exact = terms.sum(-1)
    d1 = terms[:, :, :16].sum(-1) if two else np.zeros_like(exact)
    m1 = np.abs(terms[:, :, :16]).max(-1)
    m2 = np.maximum(np.abs(terms[:, :, 16:]).max(-1), np.abs(d1)) if two else np.zeros_like(exact)
    err = np.abs(got - exact)
    rnd = 0.5 * (ulp16(d1) * two + ulp16(exact))              # round-to-nearest after each step
    resid = np.maximum(err - rnd, 0.0)
    coef = resid / np.maximum(m1 + m2, 1e-30)
    print(f"n={n} d={d} {kind}: sigma={sigma:g} max|D|={np.abs(exact).max():.0f} max addend={max(m1.max(), m2.max()):.0f} "
          f"max err/(ulp(D1)+ulp(D))={np.max(err / (ulp16(d1) * two + ulp16(exact))):.3f}   "
          f"beyond RN: max resid/(M1+M2)=2^{np.log2(max(coef.max(), 1e-30)):.2f}  frac entries beyond RN={np.mean(resid > 0):.4f}")
    return coef.max()


if __name__ == "__main__":
    worst = 0.0
    cases = [(2048, d, "normal") for d in (29, 24, 20, 17, 16, 15, 14, 13, 7)] + [(4096, 29, "traj"), (2048, 29, "offset"),
                                                                               (2048, 29, "outliers"), (2048, 14, "outliers")]
    for (n, d, kind) in cases:
        worst = max(worst, run(n, d, n + d, kind))
    print(f"WORST internal coefficient = 2^{np.log2(worst):.2f}")
