"""GPU parity of the IW entropy / KL kernels and of the reference-surface functions against the
oracle and the golden vectors generated from the unmodified reference.
Bar (north_star): entropy, KL and gradients within 1e-5 relative in fp32; the float64 kernels are
held to 1e-9."""
import math

import numpy as np
import pytest
import torch

from conftest import golden_files, load_golden

pytestmark = pytest.mark.gpu

import mepol_b200
from mepol_b200 import entropy as E
from oracle import iw_oracle as O


def _handle(g, dev):
    idx = torch.from_numpy(g["idx"].astype(np.int32)).to(dev)
    rk = torch.from_numpy(g["dist"][:, int(g["k"])].copy()).to(dev)
    return E.SearchHandle(idx, rk, int(g["k"]) + 1)


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-10), (torch.float32, 1e-5)])
@pytest.mark.parametrize("path", golden_files("iw_"), ids=lambda p: p.split("iw_")[-1][:-4])
def test_kernels_match_reference_golden(path, dtype, tol):
    g = load_golden(path)
    dev = torch.device("cuda")
    k, ns, eps = int(g["k"]), int(g["ns"]), float(g["eps"])
    lens = g["lens"].reshape(-1)
    off, n = mepol_b200.traj_offsets(lens, dev)
    lt = torch.tensor(g["logp_target"], dtype=dtype, device=dev, requires_grad=True)
    lb = torch.tensor(g["logp_behav"], dtype=dtype, device=dev)
    H, kl_raw, n_bad = mepol_b200.iw_entropy(lt, lb, _handle(g, dev), off, len(lens), k, ns,
                                             float(g["G"]), float(g["B"]), eps)
    (-H).backward()
    np.testing.assert_allclose(H.item(), float(g["entropy"]), rtol=tol)
    np.testing.assert_allclose(max(0.0, kl_raw.item()), float(g["kl"]), rtol=10 * tol, atol=1e-7 if dtype == torch.float32 else 1e-13)
    assert n_bad.item() == 0
    ref = g["grad_loss_logp"]
    got = lt.grad.double().cpu().numpy()
    np.testing.assert_allclose(got / np.abs(ref).max(), ref / np.abs(ref).max(), rtol=0, atol=tol)


def _make_policy(g, who):
    import torch.nn as nn

    class P(nn.Module):
        def __init__(self):
            super().__init__()
            self.log_std = nn.Parameter(torch.zeros(int(g["A"]), dtype=torch.float64))
            layers, prev = [], int(g["F"])
            for h in g["hidden"]:
                layers += [nn.Linear(prev, int(h)).double(), nn.ReLU()]
                prev = int(h)
            self.net = nn.Sequential(*layers)
            self.mean = nn.Linear(prev, int(g["A"])).double()
    p = P()
    p.load_state_dict({k_[4:]: torch.tensor(v) for k_, v in g.items() if k_.startswith(who + ".")})
    return p


@pytest.mark.parametrize("path", golden_files("iw_"), ids=lambda p: p.split("iw_")[-1][:-4])
def test_reference_surface_matches_golden(path):
    """compute_entropy / compute_kl / backward through the same call signature the reference's
    mepol() uses, with CPU float64 policies and CPU tensors in and out."""
    g = load_golden(path)
    beh, tgt = _make_policy(g, "beh"), _make_policy(g, "tgt")
    T = g["lens"].shape[0]
    states = torch.tensor(g["states"], dtype=torch.float64)
    actions = torch.tensor(g["actions"], dtype=torch.float64)
    lens = torch.tensor(g["lens"], dtype=torch.int64)
    dist = torch.tensor(g["dist"], dtype=torch.float64)
    idx = torch.tensor(g["idx"].astype(np.int64))
    k, ns, eps = int(g["k"]), int(g["ns"]), float(g["eps"])
    ent = mepol_b200.compute_entropy(beh, tgt, states, actions, T, lens, dist, idx, k, float(g["G"]),
                                     float(g["B"]), ns, eps)
    assert ent.device.type == "cpu" and ent.dtype == torch.float64 and ent.dim() == 0
    np.testing.assert_allclose(ent.detach().numpy(), g["entropy"], rtol=1e-10)
    (-ent).backward()
    for name, p in tgt.named_parameters():
        ref = g["grad." + name]
        sc = np.abs(ref).max() + 1e-300
        np.testing.assert_allclose(p.grad.numpy() / sc, ref / sc, rtol=0, atol=1e-9, err_msg=name)
    with torch.no_grad():
        kl, err = mepol_b200.compute_kl(beh, tgt, states, actions, T, lens, dist, idx, k, eps)
    np.testing.assert_allclose(kl.numpy(), g["kl"], rtol=1e-9, atol=1e-13)
    assert bool(err) == bool(g["kl_numeric_error"])
    w = mepol_b200.compute_importance_weights(beh, tgt, states, actions, T, lens)
    np.testing.assert_allclose(w.numpy(), g["w"], rtol=1e-9)


def test_policy_update_matches_oracle_step():
    """One optimiser step through policy_update equals an SGD step with the oracle's gradient."""
    g = load_golden(golden_files("iw_ragged")[0])
    beh, tgt = _make_policy(g, "beh"), _make_policy(g, "tgt")
    before = {n: p.detach().clone() for n, p in tgt.named_parameters()}
    opt = torch.optim.SGD(tgt.parameters(), lr=1e-3)
    T = g["lens"].shape[0]
    args = (torch.tensor(g["states"], dtype=torch.float64), torch.tensor(g["actions"], dtype=torch.float64), T,
            torch.tensor(g["lens"], dtype=torch.int64), torch.tensor(g["dist"]), torch.tensor(g["idx"].astype(np.int64)),
            int(g["k"]), float(g["G"]), float(g["B"]), int(g["ns"]), float(g["eps"]))
    loss, numeric_error = mepol_b200.policy_update(opt, beh, tgt, *args)
    assert not bool(numeric_error)
    np.testing.assert_allclose(-loss.detach().numpy(), g["entropy"], rtol=1e-10)
    for n, p in tgt.named_parameters():
        expect = before[n].numpy() - 1e-3 * g["grad." + n]
        np.testing.assert_allclose(p.detach().numpy(), expect, rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-10), (torch.float32, 1e-5)])
def test_large_synthetic_batch_matches_oracle(dtype, tol):
    """N=200k, L=500, k=4, ns=29 -- the benchmark's shape at a size the numpy oracle does in seconds."""
    rng = np.random.default_rng(17)
    T, L, k, ns = 400, 500, 4, 29
    n = T * L
    idx = rng.integers(0, n, (n, k + 1)).astype(np.int32)
    idx[:, 0] = np.arange(n)
    rk = rng.uniform(3.0, 6.0, n)
    lt = rng.normal(0, 0.01, n)
    lb = np.zeros(n)
    lens = np.full(T, L)
    G = math.gamma(ns / 2 + 1)
    B = math.log(k) - 1.2561176684318  # digamma(4)
    dist = np.zeros((n, k + 1)); dist[:, k] = rk
    H_ref = O.entropy(lt, lb, lens, dist, idx, k, G, B, ns, 0.0)
    kl_ref, _ = O.kl(lt, lb, lens, idx, k, 0.0)
    g_ref = O.entropy_grad_logp(lt, lb, lens, dist, idx, k, G, ns, 0.0)
    dev = torch.device("cuda")
    h = E.SearchHandle(torch.from_numpy(idx).to(dev), torch.from_numpy(rk).to(dev), k + 1)
    off, _ = mepol_b200.traj_offsets(lens, dev)
    ltt = torch.tensor(lt, dtype=dtype, device=dev, requires_grad=True)
    H, kl_raw, n_bad = mepol_b200.iw_entropy(ltt, torch.tensor(lb, dtype=dtype, device=dev), h, off, T, k, ns, G, B, 0.0)
    H.backward()
    np.testing.assert_allclose(H.item(), H_ref, rtol=tol)
    np.testing.assert_allclose(max(kl_raw.item(), 0.0), kl_ref, rtol=10 * tol, atol=1e-7 if dtype == torch.float32 else 1e-13)
    got = ltt.grad.double().cpu().numpy()
    np.testing.assert_allclose(got / np.abs(g_ref).max(), g_ref / np.abs(g_ref).max(), rtol=0, atol=tol)
    # determinism: bitwise identical on a second run
    ltt2 = torch.tensor(lt, dtype=dtype, device=dev, requires_grad=True)
    H2, _, _ = mepol_b200.iw_entropy(ltt2, torch.tensor(lb, dtype=dtype, device=dev), h, off, T, k, ns, G, B, 0.0)
    H2.backward()
    assert H2.item() == H.item() and torch.equal(ltt2.grad, ltt.grad)


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-10), (torch.float32, 1e-5)])
@pytest.mark.parametrize("flavour", ["short", "long", "aligned"])
def test_many_short_ragged_trajectories_use_the_warp_kernels(dtype, tol, flavour):
    """>= 512 trajectories switches the scan / backward to the warp-per-trajectory kernels (each lane owns 16
    consecutive steps, one shuffle scan per 512 steps).  "short": ragged lengths incl. 1 and > 128 (unaligned scalar
    loads, partly filled chunks); "long": trajectories beyond 512 steps (several blocks linked by the carry, incl.
    exactly 512 / 513 / 1024 / 1025); "aligned": every offset a multiple of 4 (the 16-byte vector path) with lengths
    that are and are not multiples of 16."""
    rng = np.random.default_rng(23)
    T, k, ns = 900, 4, 7
    if flavour == "short":
        lens = rng.integers(1, 300, T)
        lens[:3] = [1, 129, 257]
    elif flavour == "long":
        T = 600
        lens = rng.integers(1, 1500, T)
        lens[:6] = [512, 513, 1024, 1025, 511, 1]
    else:
        lens = 4 * rng.integers(1, 200, T)
        lens[:4] = [16, 512, 500, 4]
    n = int(lens.sum())
    idx = rng.integers(0, n, (n, k + 1)).astype(np.int32)
    idx[:, 0] = np.arange(n)
    rk = rng.uniform(0.1, 2.0, n)
    lt = rng.normal(0, 0.03, n)
    lb = rng.normal(0, 0.03, n)
    G = math.gamma(ns / 2 + 1)
    B = math.log(k) - 1.2561176684318
    dist = np.zeros((n, k + 1)); dist[:, k] = rk
    H_ref = O.entropy(lt, lb, lens, dist, idx, k, G, B, ns, 0.0)
    kl_ref, _ = O.kl(lt, lb, lens, idx, k, 0.0)
    g_ref = O.entropy_grad_logp(lt, lb, lens, dist, idx, k, G, ns, 0.0)
    dev = torch.device("cuda")
    h = E.SearchHandle(torch.from_numpy(idx).to(dev), torch.from_numpy(rk).to(dev), k + 1)
    off, _ = mepol_b200.traj_offsets(lens, dev)
    ltt = torch.tensor(lt, dtype=dtype, device=dev, requires_grad=True)
    H, kl_raw, n_bad = mepol_b200.iw_entropy(ltt, torch.tensor(lb, dtype=dtype, device=dev), h, off, T, k, ns, G, B, 0.0)
    H.backward()
    np.testing.assert_allclose(H.item(), H_ref, rtol=tol)
    np.testing.assert_allclose(max(kl_raw.item(), 0.0), kl_ref, rtol=10 * tol, atol=1e-7 if dtype == torch.float32 else 1e-13)
    got = ltt.grad.double().cpu().numpy()
    np.testing.assert_allclose(got / np.abs(g_ref).max(), g_ref / np.abs(g_ref).max(), rtol=0, atol=tol)


def test_numeric_error_flags():
    """exp overflow (float64 range) and zero radius ('knn aliasing') give nan/inf exactly where the
    reference does (mepol.py:168, :276, :485)."""
    dev = torch.device("cuda")
    n, k = 64, 4
    idx = torch.stack([(torch.arange(n) + j) % n for j in range(k + 1)], 1).to(torch.int32).to(dev)
    off, _ = mepol_b200.traj_offsets([n], dev)
    rk = torch.ones(n, dtype=torch.float64, device=dev)
    lt = torch.zeros(n, dtype=torch.float64, device=dev)
    lt[3] = 800.0
    H, kl_raw, _ = mepol_b200.iw_entropy(lt, torch.zeros_like(lt), E.SearchHandle(idx, rk, k + 1), off, 1, k, 3, None, 0.0, 0.0)
    assert math.isnan(H.item()) and math.isnan(kl_raw.item())
    rk0 = rk.clone(); rk0[5] = 0.0
    H, kl_raw, n_bad = mepol_b200.iw_entropy(torch.zeros_like(lt), torch.zeros_like(lt), E.SearchHandle(idx, rk0, k + 1),
                                             off, 1, k, 3, None, 0.0, 0.0)
    assert math.isinf(H.item()) and H.item() < 0 and n_bad.item() == 1 and math.isfinite(kl_raw.item())


def test_rev_csr_is_the_transpose():
    rng = np.random.default_rng(4)
    n, k = 5000, 4
    idx = rng.integers(0, n, (n, k + 1)).astype(np.int32)
    h = E.SearchHandle(torch.from_numpy(idx).cuda(), torch.ones(n, dtype=torch.float64).cuda(), k + 1)
    rowptr, col = [t.cpu().numpy() for t in h.csr()]
    assert rowptr[0] == 0 and rowptr[-1] == n * k
    for j in rng.choice(n, 200, replace=False):
        expect = np.sort(np.nonzero((idx[:, :k] == j).any(1))[0].repeat(1))
        rows = np.repeat(np.arange(n), k)[idx[:, :k].reshape(-1) == j]
        np.testing.assert_array_equal(col[rowptr[j]:rowptr[j + 1]], np.sort(rows))


def test_one_graph_serves_new_batches_and_handles_of_the_same_shape():
    """A captured objective step is keyed by shape: a new epoch's batch and search result of equal shape
    are copied into the graph's buffers and replayed.  Results must equal the eager path's."""
    import torch.nn as nn
    T, L, F, A, k, ns = 6, 40, 5, 2, 3, 5
    n = T * L
    G_, B_ = math.gamma(ns / 2 + 1), 0.1

    class P(nn.Module):                                  # GaussianPolicy-shaped (src/policy.py:17-29)
        def __init__(self, seed):
            super().__init__()
            torch.manual_seed(seed)
            self.log_std = nn.Parameter(-0.5 * torch.ones(A, dtype=torch.float64))
            self.net = nn.Sequential(nn.Linear(F, 16).double(), nn.ReLU(), nn.Linear(16, 12).double(), nn.ReLU())
            self.mean = nn.Linear(12, A).double()
    beh, tgt = P(1), P(2)
    out = {}
    try:
        for graphs in (True, False):
            mepol_b200.configure(cuda_graphs=graphs)
            E._batches.clear(); E._handles.clear(); E._graphs.clear(); E._graph_seen.clear()
            vals = []
            for epoch in range(3):                       # three different batches / searches, one shape
                ge = torch.Generator().manual_seed(100 + epoch)
                states = torch.randn((T, L + 1, F), generator=ge, dtype=torch.float64)
                actions = torch.randn((T, L, A), generator=ge, dtype=torch.float64)
                lens = torch.full((T, 1), L, dtype=torch.int64)
                x = states[:, 1:].reshape(n, F).numpy().astype(np.float32)
                dist, idx = mepol_b200.knn(x, k)
                dist, idx = torch.from_numpy(dist), torch.from_numpy(idx)
                for _ in range(3):
                    for p in tgt.parameters():
                        p.grad = None
                    H = mepol_b200.compute_entropy(beh, tgt, states, actions, T, lens, dist, idx, k, G_, B_, ns, 0.0)
                    H.backward()
                    vals.append((H.detach().clone(), [p.grad.clone() for p in tgt.parameters()]))
            out[graphs] = vals
            if graphs:
                assert len(E._graphs) == 1, "one shape -> one captured graph"
    finally:
        mepol_b200.configure(cuda_graphs=True)
    for (h1, g1), (h2, g2) in zip(out[True], out[False]):
        assert torch.equal(h1, h2)
        for a, b in zip(g1, g2):
            assert torch.equal(a, b)


@pytest.mark.parametrize("n_dup", [3000, 200000])
def test_rev_csr_hub_rows_from_exact_duplicates(n_dup):
    """In-degree is not bounded by k: with exact duplicates every copy lists the same k lowest-index
    copies, so k hub rows collect ~n_dup entries each.  The long rows go through the block
    radix sort: still the sorted transpose, and in milliseconds, not minutes."""
    import time
    rng = np.random.default_rng(9)
    n, k = n_dup + 4000, 4
    idx = rng.integers(0, n, (n, k + 1)).astype(np.int32)
    idx[:n_dup, :k] = np.array([7, 11, 13, 17], dtype=np.int32)      # every duplicate points at the same hubs
    idx[:n_dup, 0] = np.where(np.arange(n_dup) % 3 == 0, 7, 19)       # a second, shorter hub
    h = E.SearchHandle(torch.from_numpy(idx).cuda(), torch.ones(n, dtype=torch.float64).cuda(), k + 1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rowptr, col = h.csr()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    rowptr, col = rowptr.cpu().numpy(), col.cpu().numpy()
    assert rowptr[0] == 0 and rowptr[-1] == n * k
    flat = idx[:, :k].reshape(-1)
    src = np.repeat(np.arange(n), k)
    for j in [7, 11, 13, 17, 19] + list(rng.choice(n, 50, replace=False)):
        np.testing.assert_array_equal(col[rowptr[j]:rowptr[j + 1]], np.sort(src[flat == j]))
    assert dt < 1.0, f"reverse CSR with {n_dup}-entry hub rows took {dt:.2f} s"
    # and the backward that walks those rows matches the oracle
    from oracle import iw_oracle
    lt = rng.normal(0, 0.02, n)
    lens = np.array([n])
    rk = np.abs(rng.normal(1.0, 0.1, n)) + 0.5
    dist = np.zeros((n, k + 1)); dist[:, k] = rk
    g_ref = iw_oracle.entropy_grad_logp(lt, np.zeros(n), lens, dist, idx.astype(np.int64), k, math.gamma(2.5), 3, 0.0)
    h2 = E.SearchHandle(torch.from_numpy(idx).cuda(), torch.from_numpy(rk).cuda(), k + 1)
    off, _ = mepol_b200.traj_offsets(lens, "cuda")
    ltt = torch.tensor(lt, dtype=torch.float64, device="cuda", requires_grad=True)
    H, _, _ = mepol_b200.iw_entropy(ltt, torch.zeros_like(ltt), h2, off, 1, k, 3, math.gamma(2.5), 0.0, 0.0)
    H.backward()
    got = ltt.grad.cpu().numpy()
    assert np.max(np.abs(got - g_ref)) <= 1e-9 * np.max(np.abs(g_ref))


def test_cuda_graph_replay_equals_eager_launches_over_several_steps():
    """The captured objective step (default) and the eager launches run the same kernels: losses,
    KLs and parameters after several optimiser steps are bit-identical, and the RNG burn leaves the
    CPU generator in the same state either way."""
    from mepol_b200 import entropy as E
    g = load_golden(golden_files("iw_ragged")[0])
    T = g["lens"].shape[0]
    args = (torch.tensor(g["states"], dtype=torch.float64), torch.tensor(g["actions"], dtype=torch.float64), T,
            torch.tensor(g["lens"], dtype=torch.int64), torch.tensor(g["dist"]), torch.tensor(g["idx"].astype(np.int64)),
            int(g["k"]), float(g["G"]), float(g["B"]), int(g["ns"]), float(g["eps"]))
    runs = {}
    try:
        for graphs in (True, False):
            mepol_b200.configure(cuda_graphs=graphs)
            E._batches.clear()
            E._handles.clear()
            E._graphs.clear()
            E._graph_seen.clear()
            beh, tgt = _make_policy(g, "beh"), _make_policy(g, "tgt")
            opt = torch.optim.Adam(tgt.parameters(), lr=1e-3)
            torch.manual_seed(7)
            trace = []
            for _ in range(4):
                loss, bad = mepol_b200.policy_update(opt, beh, tgt, *args)
                with torch.no_grad():
                    kl, err = mepol_b200.compute_kl(beh, tgt, *args[:7], args[10])
                trace.append((loss.detach().clone(), kl.clone()))
            runs[graphs] = (trace, [p.detach().clone() for p in tgt.parameters()], torch.get_rng_state().clone())
            if graphs:
                n_captured = len(E._graphs)
    finally:
        mepol_b200.configure(cuda_graphs=True)
    assert E._graph_failures == 0 and n_captured >= 1      # steps 2.. of the graph run were replays
    for (l1, k1), (l2, k2) in zip(runs[True][0], runs[False][0]):
        assert torch.equal(l1, l2) and torch.equal(k1, k2)
    for a, b in zip(runs[True][1], runs[False][1]):
        assert torch.equal(a, b)
    assert torch.equal(runs[True][2], runs[False][2])


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-12), (torch.float32, 2e-5)])
@pytest.mark.parametrize("n,A", [(10000, 8), (777, 1), (24000, 2), (5000, 17)])
def test_gauss_logp_kernels_match_the_reference_expression(n, A, dtype, tol):
    """mepol_gauss_logp_{fwd,bwd} against autograd of the literal expression of policy.py:43-51
    (float64 on the CPU): value, d/d mean and d/d log_std (a reduction over all states)."""
    g = torch.Generator().manual_seed(n + A)
    mean = torch.randn((n, A), generator=g, dtype=torch.float64)
    act = 0.6 * torch.randn((n, A), generator=g, dtype=torch.float64)
    ls = -0.5 + 0.3 * torch.randn(A, generator=g, dtype=torch.float64)
    up = torch.randn(n, generator=g, dtype=torch.float64)
    m_ref, ls_ref = mean.clone().requires_grad_(True), ls.clone().requires_grad_(True)
    ref = torch.sum(-0.5 * (math.log(2 * math.pi) + 2 * ls_ref + (act - m_ref) ** 2 / (torch.exp(ls_ref) + 1e-7) ** 2), dim=1)
    ref.backward(up)
    m_dev = mean.to("cuda", dtype).requires_grad_(True)
    ls_dev = ls.to("cuda", dtype).requires_grad_(True)
    out = E.GaussLogPFunction.apply(m_dev, act.to("cuda", dtype), ls_dev)
    out.backward(up.to("cuda", dtype))
    sc = ref.detach().abs().max().item()
    assert (out.detach().cpu().double() - ref.detach()).abs().max().item() <= tol * sc
    sc = m_ref.grad.abs().max().item()
    assert (m_dev.grad.cpu().double() - m_ref.grad).abs().max().item() <= tol * sc
    sc = ls_ref.grad.abs().max().item()
    assert (ls_dev.grad.cpu().double() - ls_ref.grad).abs().max().item() <= 10 * tol * sc
    # deterministic: the log_std reduction folds block partials in a fixed order
    m2 = mean.to("cuda", dtype).requires_grad_(True)
    ls2 = ls.to("cuda", dtype).requires_grad_(True)
    E.GaussLogPFunction.apply(m2, act.to("cuda", dtype), ls2).backward(up.to("cuda", dtype))
    assert torch.equal(ls2.grad, ls_dev.grad) and torch.equal(m2.grad, m_dev.grad)
