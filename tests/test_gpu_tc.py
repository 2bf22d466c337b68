"""GPU tests specific to the tcgen05 pruner: the raw tensor-core score tile against a float64
emulation of the fp16 operands, and bit-identity of the full search with the SIMT path (both are
exact, so they must agree on every index and every distance)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import mepol_b200
from mepol_b200 import _lib
from conftest import assert_knn_matches


def _expected_scores(x, nq=256, nr=128):
    """Mirror of knn_prep.cu's packing: centre with the float32-rounded mean, scale by the power of
    two that brings the column-extent bound of the max norm under 128, round to fp16, split ||y'||^2
    over three fp16 terms; D = sum(n_hi, n_mid, n_lo) - 2 x~.y~ in float64."""
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
    return ((h + m + l)[:nr][None, :] - 2.0 * (xb[:nq] @ xb[:nr].T)).numpy()


@pytest.mark.parametrize("d", [29, 24, 7, 2])
def test_tensor_core_score_tile(d):
    rng = np.random.default_rng(d)
    n = 1024
    x = rng.standard_normal((n, d)).astype(np.float32)
    xt = torch.from_numpy(x).cuda()
    out = torch.full((256, 128), float("nan"), dtype=torch.float32, device="cuda")
    ws = torch.empty(_lib.knn_workspace_bytes(n, n, d, 5, _lib.KNN_TC), dtype=torch.uint8, device="cuda")
    rc = _lib.load().mepol_debug_tc_scores(xt.data_ptr(), n, xt.data_ptr(), n, d, 5, out.data_ptr(), ws.data_ptr(),
                                           ws.numel(), torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "mepol_debug_tc_scores")
    got = out.cpu().numpy().astype(np.float64)
    exp = _expected_scores(x)
    assert not np.isnan(got).any()
    # fp32 accumulation of 32 exact fp16 products: a few ulp of the largest term
    assert np.max(np.abs(got - exp)) < 1e-4 * max(1.0, np.abs(exp).max())


@pytest.mark.parametrize("n,d,k1", [(70000, 29, 5), (33000, 29, 2), (20000, 16, 5), (9000, 29, 12), (40000, 7, 5)])
def test_tc_and_simt_searches_are_bit_identical(n, d, k1):
    rng = np.random.default_rng(n)
    x = torch.from_numpy(rng.standard_normal((n, d)).astype(np.float32)).cuda()
    a = mepol_b200.knn_device(x, k1, algo=_lib.KNN_SIMT)
    b = mepol_b200.knn_device(x, k1, algo=_lib.KNN_TC)
    assert torch.equal(a.indices, b.indices)
    assert torch.equal(a.distances, b.distances)
    assert torch.equal(a.rk, b.rk)


def test_tc_rejects_unsupported_shapes():
    x = torch.zeros((5000, 40), device="cuda")
    with pytest.raises(_lib.MepolError):
        mepol_b200.knn_device(x, 5, algo=_lib.KNN_TC)         # d + 3 > 32
    x = torch.zeros((5000, 8), device="cuda")
    with pytest.raises(_lib.MepolError):
        mepol_b200.knn_device(x, 30, algo=_lib.KNN_TC)        # k1 too large for the 32-entry soft cap
    res = mepol_b200.knn_device(torch.randn((5000, 40), device="cuda"), 5)   # AUTO falls back to SIMT
    assert res.indices.shape == (5000, 5)


def test_tc_offset_clustered_data_uses_band_and_fallback_correctly():
    """Tight clusters far from the origin: after centring the fp16 band is wide relative to the
    neighbour gaps, so many rows overflow into the exact fallback -- results must still be exact."""
    from oracle import knn_oracle
    rng = np.random.default_rng(8)
    centers = rng.standard_normal((20, 29)) * 30
    x = (centers[rng.integers(0, 20, 12000)] + 0.01 * rng.standard_normal((12000, 29))).astype(np.float32)
    ref_d, ref_i = knn_oracle.knn_c(x, 5, mode="kd_tree")
    res = mepol_b200.knn_device(torch.from_numpy(x).cuda(), 5, algo=_lib.KNN_TC)
    np.testing.assert_array_equal(res.indices.cpu().numpy(), ref_i)
    np.testing.assert_allclose(res.distances.cpu().numpy(), ref_d, rtol=1e-12)


def test_tc_queries_outside_fp16_range_take_the_exact_fallback():
    """Queries far outside the fitted set's extent cannot be represented in the scaled fp16 operand;
    they get an infinite band, overflow, and are answered by the exact fallback."""
    from oracle import knn_oracle
    rng = np.random.default_rng(11)
    x = rng.standard_normal((6000, 29)).astype(np.float32)
    q = rng.standard_normal((300, 29)).astype(np.float32)
    q[:7] *= 1e7
    q[7] += 3e5
    res = mepol_b200.knn_device(torch.from_numpy(x).cuda(), 5, queries=torch.from_numpy(q).cuda(), algo=_lib.KNN_TC)
    ref_d, ref_i = knn_oracle.knn_c(x, 5, queries=q, mode="kd_tree")
    assert res.stats_dict()["tc_overflow_rows"] >= 8
    assert_knn_matches(res.distances.cpu().numpy(), res.indices.cpu().numpy(), ref_d, ref_i, x, queries=q)


@pytest.mark.parametrize("scale", [1e-20, 1e-3, 1e4, 1e15])
def test_tc_is_scale_invariant(scale):
    """The power-of-two operand scaling keeps the fp16 pruner usable at any data magnitude."""
    rng = np.random.default_rng(5)
    x = (rng.standard_normal((20000, 29)) * scale).astype(np.float32)
    xt = torch.from_numpy(x).cuda()
    a = mepol_b200.knn_device(xt, 5, algo=_lib.KNN_SIMT)
    b = mepol_b200.knn_device(xt, 5, algo=_lib.KNN_TC)
    assert torch.equal(a.indices, b.indices)
    assert torch.equal(a.distances, b.distances)
    assert b.stats_dict()["tc_overflow_rows"] == 0


@pytest.mark.parametrize("n_ref,n_q", [(70000, 50688), (140000, 45568)])   # 2 and 4 reference splits on 148 SMs
def test_tc_tail_round_reference_splits_match_simt(n_ref, n_q):
    """Query-row counts whose last persistent-grid round is partial (the multi-GPU shard shape): the
    tail row blocks are cut into reference-range splits with separate candidate lists that the
    re-rank merges -- results must not change."""
    rng = np.random.default_rng(n_q)
    x = torch.from_numpy(rng.standard_normal((n_ref, 29)).astype(np.float32)).cuda()
    q = x[:n_q]
    a = mepol_b200.knn_device(x, 5, queries=q, algo=_lib.KNN_TC)
    b = mepol_b200.knn_device(x, 5, queries=q, algo=_lib.KNN_SIMT)
    assert torch.equal(a.indices, b.indices)
    assert torch.equal(a.distances, b.distances)
    qq = torch.from_numpy(rng.standard_normal((n_q, 29)).astype(np.float32)).cuda()       # not a slice of the fitted set
    a = mepol_b200.knn_device(x, 5, queries=qq, algo=_lib.KNN_TC)
    b = mepol_b200.knn_device(x, 5, queries=qq, algo=_lib.KNN_SIMT)
    assert torch.equal(a.indices, b.indices)
    assert torch.equal(a.distances, b.distances)


def test_tc_mass_overflow_goes_through_the_fp32_rescue_pass():
    """Forcing the tensor-core pruner onto 2-D data overflows the fp16 band for almost every row; past
    16384 flagged rows the SIMT rescue pass re-prunes them in place instead of the per-row exact
    fallback.  Results stay exact."""
    rng = np.random.default_rng(21)
    x = torch.from_numpy(rng.uniform(-1, 1, (40000, 2)).astype(np.float32)).cuda()
    a = mepol_b200.knn_device(x, 5, algo=_lib.KNN_TC)
    b = mepol_b200.knn_device(x, 5, algo=_lib.KNN_GRID)
    assert a.stats_dict()["tc_overflow_rows"] > 16384
    assert torch.equal(a.indices, b.indices)
    assert torch.equal(a.distances, b.distances)


# ------------------------------------------------------------------------------------------------
# Second-generation pruner (knn_tc2.cu): N = 256 MMAs with fp16 accumulators.  The band the kernel
# prunes with assumes at most one fp16 ulp of error per accumulation step;
# the raw tile pins what the hardware really does.
# ------------------------------------------------------------------------------------------------
def _ulp16(v):
    a = np.maximum(np.abs(v), 2.0 ** -14)
    return 2.0 ** (np.floor(np.log2(a)) - 10)


def _tc2_operands(x):
    """knn_prep.cu's packing in float64 with correctly rounded fp16 conversions (numpy rounds
    float64 -> float16 directly; torch goes through float32 and double-rounds near ties)."""
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
    return xb, [h, m, l]


@pytest.mark.parametrize("d,kind", [(29, "normal"), (29, "traj"), (29, "offset"), (24, "normal"), (16, "normal"),
                                    (14, "normal"), (13, "normal"), (7, "normal")])
def test_tc2_fp16_accumulator_tile_and_rounding_model(d, kind):
    rng = np.random.default_rng(100 + d)
    n = 2048
    x = rng.standard_normal((n, d))
    if kind == "traj":
        x = np.cumsum(0.05 * x, axis=0)
    if kind == "offset":
        x = x + 100.0
    x = x.astype(np.float32)
    xt = torch.from_numpy(x).cuda()
    out = torch.full((256, 256), float("nan"), dtype=torch.float32, device="cuda")
    ws = torch.empty(_lib.knn_workspace_bytes(n, n, d, 5, _lib.KNN_TC), dtype=torch.uint8, device="cuda")
    rc = _lib.load().mepol_debug_tc2_scores(xt.data_ptr(), n, xt.data_ptr(), n, d, 5, out.data_ptr(), ws.data_ptr(),
                                            ws.numel(), torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "mepol_debug_tc2_scores")
    got = out.cpu().numpy().astype(np.float64)
    assert not np.isnan(got).any()
    xb, nrm = _tc2_operands(x)
    q, r = xb[:256], xb[:256]
    # operand columns: d coordinates, then n_hi, n_mid, n_lo; the first K = 16 step covers columns 0..15
    cols_q = np.concatenate([q, np.ones((256, 3))], axis=1)
    cols_r = np.concatenate([-2.0 * r, np.stack([nv[:256] for nv in nrm], axis=1)], axis=1)
    terms = cols_q[:, None, :] * cols_r[None, :, :]             # exact products (fp16 x fp16 fits float64)
    exact = terms.sum(-1)
    two_steps = d + 3 > 16
    d1 = terms[:, :, :16].sum(-1) if two_steps else np.zeros_like(exact)
    m1 = np.abs(terms[:, :, :16]).max(-1)
    m2 = np.maximum(np.abs(terms[:, :, 16:]).max(-1), np.abs(d1)) if two_steps else np.zeros_like(exact)
    err = np.abs(got - exact)
    # Measured on B200 (scripts/exp_tc2.py): each K = 16 step forms its 17-addend sum at fp32-like precision
    # (<= 2^-25 of the largest addend) and stores it rounded to nearest fp16 (<= 0.5 ulp).  The kernel's band
    # assumes 1 ulp per step and 2^-20 of the largest addend; this test fails well before those are reached.
    allowed = 0.5 * ((_ulp16(d1) if two_steps else 0.0) + _ulp16(exact)) + 2.0 ** -22 * (m1 + m2)
    assert np.all(err <= allowed), float(np.max(err / allowed))


@pytest.mark.parametrize("variant", ["0", "2", "3", "4", "5"])
def test_every_tensor_core_variant_is_bit_identical_to_simt(variant, monkeypatch):
    """MEPOL_TC_VARIANT is an A/B switch between exact kernels (csrc/capi.cu::tc_variant)."""
    rng = np.random.default_rng(77)
    x = torch.from_numpy(rng.standard_normal((70000, 29)).astype(np.float32)).cuda()
    a = mepol_b200.knn_device(x, 5, algo=_lib.KNN_SIMT)
    monkeypatch.setenv("MEPOL_TC_VARIANT", variant)
    b = mepol_b200.knn_device(x, 5, algo=_lib.KNN_TC)
    assert torch.equal(a.indices, b.indices)
    assert torch.equal(a.distances, b.distances)
    q = x[:45568]                                    # partial last round -> reference-range splits
    a = mepol_b200.knn_device(x, 5, queries=q, algo=_lib.KNN_SIMT)
    b = mepol_b200.knn_device(x, 5, queries=q, algo=_lib.KNN_TC)
    assert torch.equal(a.indices, b.indices)
    assert torch.equal(a.distances, b.distances)


def _adversarial(kind, rng):
    n, d = 30000, 29
    x = rng.standard_normal((n, d))
    if kind == "huge_column":          # one coordinate dominates every norm: massive cancellation in the scores
        x[:, 0] *= 1000.0
    elif kind == "scaling_edge":       # an outlier that lands exactly on the fp16 scaling limit (norm 128 after scaling)
        x[0] = 0.0
        x[0, 3] = np.abs(x).max() * 7.0
    elif kind == "outliers":           # a few far points with huge ||y||^2 next to a dense core
        x[:50] *= 12.0
    elif kind == "tight_trajectories":  # neighbour gaps far below the row norms
        x = np.cumsum(0.01 * x.reshape(60, 500, d), axis=1).reshape(n, d) + rng.standard_normal((60, 1, d)).repeat(500, 1).reshape(n, d) * 3
    elif kind == "first_half_heavy":   # all the energy in the first 16 coordinates (the first K step)
        x[:, 16:] *= 1e-3
    elif kind == "second_half_heavy":
        x[:, :16] *= 1e-3
    elif kind == "lattice":            # masses of exact ties
        x = rng.integers(-3, 4, (n, d)).astype(np.float64)
    return x.astype(np.float32)


@pytest.mark.parametrize("kind", ["huge_column", "scaling_edge", "outliers", "tight_trajectories", "first_half_heavy",
                                  "second_half_heavy", "lattice"])
@pytest.mark.parametrize("variant", ["3", "2"])
def test_band_holds_on_adversarial_data(kind, variant, monkeypatch):
    """The pruning band is a proof obligation, not an observation: on data built to stress it (fp16
    scaling edge, one huge column, outliers, cancellation inside one K step, exact ties) the tensor-core
    result must still equal the fp32 SIMT pruner's bit for bit (both are exact)."""
    x = torch.from_numpy(_adversarial(kind, np.random.default_rng(5))).cuda()
    a = mepol_b200.knn_device(x, 5, algo=_lib.KNN_SIMT)
    monkeypatch.setenv("MEPOL_TC_VARIANT", variant)
    b = mepol_b200.knn_device(x, 5, algo=_lib.KNN_TC)
    assert torch.equal(a.distances, b.distances)
    if kind == "lattice":
        # exact ties: indices are decided by (d^2, index) in both paths
        assert torch.equal(a.indices, b.indices)
    else:
        assert torch.equal(a.indices, b.indices)


@pytest.mark.parametrize("factor", [3.0, 20.0, 60.0, 400.0])
@pytest.mark.parametrize("variant", ["3", "2"])
def test_foreign_queries_with_large_norms_stay_exact(factor, variant, monkeypatch):
    """Queries that are NOT rows of the fitted set may have norms far beyond its extent while every
    component still fits fp16.  With fp16 accumulators a partial sum of such a row could overflow to
    +-inf; those rows must be detected (|partial| <= 2 ||x~|| max||y~|| + A^2 >= fp16 range) and served by
    the exact fallback, the others pruned as usual -- results equal the oracle either way."""
    from oracle import knn_oracle
    rng = np.random.default_rng(int(factor))
    x = rng.standard_normal((8000, 29)).astype(np.float32)
    q = (rng.standard_normal((700, 29)) * factor).astype(np.float32)
    q[:50] = x[:50]                                       # some ordinary rows in the same batch
    monkeypatch.setenv("MEPOL_TC_VARIANT", variant)
    res = mepol_b200.knn_device(torch.from_numpy(x).cuda(), 5, queries=torch.from_numpy(q).cuda(), algo=_lib.KNN_TC)
    ref_d, ref_i = knn_oracle.knn_c(x, 5, queries=q, mode="kd_tree")
    assert_knn_matches(res.distances.cpu().numpy(), res.indices.cpu().numpy(), ref_d, ref_i, x, queries=q, rtol=1e-12)


@pytest.mark.parametrize("n,clump", [(30011, 2), (30011, 6), (50021, 40), (66000, 3)])
def test_survivor_path_branches_agree_with_simt(n, clump):
    """The pruner's survivor handling has two routes: the usual one (every row of the warp has at most ONE
    survivor in the tile: its value is the group minimum, nothing is parked) and the general one (parked
    group + walk).  Index-adjacent near-duplicates put a row's nearest neighbours into the SAME 32-score
    group of the SAME tile, clumps wider than 32 spread them over two groups, and n is not a multiple of
    32 / 256 (ragged last group, padded last tile): every combination must give the exact answer."""
    rng = np.random.default_rng(n + clump)
    base = rng.standard_normal(((n + clump - 1) // clump, 29))
    x = np.repeat(base, clump, axis=0)[:n] + 1e-3 * rng.standard_normal((n, 29))
    x = torch.from_numpy(x.astype(np.float32)).cuda()
    a = mepol_b200.knn_device(x, 5, algo=_lib.KNN_SIMT)
    b = mepol_b200.knn_device(x, 5, algo=_lib.KNN_TC)
    assert torch.equal(a.indices, b.indices)
    assert torch.equal(a.distances, b.distances)
    assert torch.equal(a.rk, b.rk)
    if clump == 3 and n == 66000:        # and against the independent oracle on a sample of rows
        from oracle import knn_oracle
        rows = np.random.default_rng(0).choice(n, 300, replace=False)
        xn = x.cpu().numpy()
        ref_d, ref_i = knn_oracle.knn_c(xn, 5, queries=xn[rows], mode="kd_tree")
        assert_knn_matches(b.distances.cpu().numpy()[rows], b.indices.cpu().numpy()[rows], ref_d, ref_i, xn,
                           queries=xn[rows], rtol=1e-12)
