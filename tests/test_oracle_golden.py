"""Pins the CPU oracle (oracle/fw_oracle.cpp) to the reference's own fixtures.

Fixtures: tests/golden/* (generated from /root/reference/test/data by
tests/golden/make_golden.py) plus the inline known answers of
/root/reference/test/statfuns.jl and /root/reference/test/contingency.jl.
"""
import json
import os

import numpy as np
import pytest
from scipy import stats as sps

from oracle import fwo


@pytest.fixture(scope="module")
def inputs(golden_dir):
    return np.load(os.path.join(golden_dir, "hmp_inputs.npz"))


@pytest.fixture(scope="module")
def expected(golden_dir):
    with open(os.path.join(golden_dir, "tests_expected.json")) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def graphs(golden_dir):
    with open(os.path.join(golden_dir, "learning_expected.json")) as f:
        return json.load(f)


def _check(got, want, atol_stat, atol_p):
    got = [got] if isinstance(got, tuple) else got
    assert len(got) == len(want)
    for g, w in zip(got, want):
        # reference tolerance is rtol 1e-2 (test/tests.jl:12-14); we hold the oracle far tighter
        assert abs(g[0] - w[0]) <= atol_stat, (g, w)
        assert abs(g[1] - w[1]) <= atol_p, (g, w)
        assert g[2] == w[2] and g[3] == w[3], (g, w)


# ---- test/tests.jl:41-74: 204 golden TestResults ----------------------------------------
@pytest.mark.parametrize("kind", ["mi", "mi_nz", "fz", "fz_nz"])
def test_golden_testresults(kind, inputs, expected):
    disc = kind.startswith("mi")
    # discrete: exact arithmetic up to log/chi2 rounding; continuous: the input TSVs are
    # Float32-rounded text and pcor_rec rounds numerators to 5 digits while the golden
    # conditional values are the exact pcor -> <= 1e-5 / 4e-5.
    tol_uni = (1e-14, 1e-13) if disc else (2e-7, 5e-7)
    tol_cond = (1e-14, 1e-13) if disc else (1e-5, 5e-5)
    o = fwo.Oracle(inputs[kind], kind, cont32=False)     # prec=64 in test/tests.jl:7-10
    if kind == "fz":
        o.compute_cor()
    _check(o.test_uni(0, list(range(1, 50))), expected[f"exp_uni_{kind}"], *tol_uni)
    mk1, mk3 = (1, 3) if disc else (3, 3)
    _check(o.test_cond(30, 20, [6], max_k=mk1), expected[f"exp_condZ1_{kind}"], *tol_cond)
    _check(o.test_cond(30, 20, [6, 13, 17], max_k=mk3), expected[f"exp_condZ3_{kind}"], *tol_cond)


# ---- test/statfuns.jl:27-40 ---------------------------------------------------------------
def test_pcor_known_answers(inputs):
    o = fwo.Oracle(inputs["fz"], "fz", cont32=False)
    cor = o.compute_cor()
    assert abs(fwo.pcor_rec(cor, 0, 15, [40], cont32=False) - (-0.16393307352649356)) < 1e-4
    assert abs(fwo.pcor_rec(cor, 30, 20, [6, 13, 17], cont32=False) - (-0.07643814205965811)) < 1e-4
    assert fwo.fz_pval(-0.16393307352649356, 351, 1) == pytest.approx(0.0020593283914246987, rel=1e-6)
    assert fwo.fz_pval(-0.07643814205965811, 351, 3) == pytest.approx(0.1548665431407692, rel=1e-6)


def test_pcor_rec_float32_promotions():
    """statfuns.jl:39-53: k=1 is all-Float32, `^2.0` promotes at k>=2; numerators rounded to 5 digits."""
    rng = np.random.default_rng(3)
    A = rng.standard_normal((200, 6))
    cor = np.corrcoef(A, rowvar=False).astype(np.float32).astype(np.float64)
    f = np.float32
    pXY, pXZ, pYZ = f(cor[0, 1]), f(cor[0, 2]), f(cor[1, 2])
    e = f(pXY - f(pXZ * pYZ))
    e = f(np.rint(f(e * f(1e5))) / f(1e5))
    d = f(np.sqrt(f(f(1) - f(pXZ * pXZ))) * np.sqrt(f(f(1) - f(pYZ * pYZ))))
    want = f(e / d)
    got = fwo.pcor_rec(cor, 0, 1, [2], cont32=True)
    assert got == float(want)
    # k=2 result is Float64-valued and depends on argument order only through b^2 (f32) vs c^2.0 (f64)
    g1 = fwo.pcor_rec(cor, 0, 1, [2, 3], cont32=True)
    assert g1 != float(np.float32(g1)) or g1 == 0.0
    exact = np.linalg.inv(cor[np.ix_([0, 1, 2, 3], [0, 1, 2, 3])])
    assert abs(g1 - (-exact[0, 1] / np.sqrt(exact[0, 0] * exact[1, 1]))) < 5e-5


# ---- test/statfuns.jl:46-57 ---------------------------------------------------------------
def test_mutual_information_known_answers():
    assert abs(fwo.mutual_information([[4, 2], [2, 4]])) == pytest.approx(0.05663301226513242, rel=1e-12)
    t3 = np.zeros((2, 2, 3), int)
    t3[0, 0, 0], t3[1, 0, 0], t3[0, 1, 1], t3[1, 1, 1], t3[1, 1, 2] = 4, 2, 2, 3, 1
    assert fwo.mutual_information(t3) == pytest.approx(0.0, abs=1e-12)
    t6 = np.zeros((2, 2, 6), int)
    t6[0, 0, 0] = 2; t6[0, 1, 1] = 2; t6[1, 1, 1] = 2; t6[0, 0, 2] = 2; t6[1, 0, 2] = 2; t6[1, 1, 3] = 1; t6[1, 1, 4] = 1
    assert fwo.mutual_information(t6) == pytest.approx(0.0, abs=1e-12)
    assert fwo.mi_pval(0.05663301226513242, 1, 351) == pytest.approx(2.8770005665168745e-10, rel=1e-6)


def test_chisq_sf_matches_scipy():
    for df in [1, 2, 3, 4, 7, 8, 27, 54, 108]:
        for x in [1e-3, 0.5, 1.0, 5.0, 20.0, 80.0, 300.0, 1500.0]:
            want = sps.chi2.sf(x, df)
            got = fwo.chisq_sf(df, x)
            assert got == pytest.approx(want, rel=2e-12, abs=1e-300), (df, x)


def test_fz_pval_matches_scipy():
    for r in [-0.9, -0.3, 0.0, 1e-4, 0.2, 0.77, 0.999]:
        for n in [4, 20, 346, 10000]:
            z = np.sqrt(n - 3) / 2 * np.log((1 + r) / (1 - r))
            assert fwo.fz_pval(r, n) == pytest.approx(2 * sps.norm.sf(abs(z)), rel=1e-12, abs=1e-300)
    assert fwo.fz_pval(0.5, 3) == 1.0        # sample_factor <= 0 -> z = 0
    assert fwo.fz_pval(1.0, 100) == 0.0


# ---- test/statfuns.jl:61-70 ---------------------------------------------------------------
def test_benjamini_hochberg_known_answer():
    pv = [0.0, 1.0, 0.973774, 0.722245, 0.805758, 0.713164, 0.314595, 0.947966, 0.001, 0.0339692]
    fdr = np.array([0.0, 1.0, 1.0, 1.0, 1.0, 1.0, 0.786488, 1.0, 0.005, 0.113231])
    got = fwo.benjamini_hochberg(pv, alpha=0.01)
    sig = got < 0.01
    assert (sig == (fdr < 0.01)).all()
    assert np.allclose(got[sig], fdr[sig], rtol=1e-6)
    assert np.isnan(got[~sig]).all()          # statfuns.jl:346: everything else becomes NaN


# ---- test/contingency.jl:5-66 -------------------------------------------------------------
def test_contingency_known_answers():
    v1 = [0, 0, 0, 0, 1, 1, 1, 1, 0, 1, 0, 1]
    v2 = [0, 0, 1, 1, 1, 1, 0, 0, 0, 1, 0, 1]
    v3 = [0, 0, 1, 1, 1, 1, 0, 0, 0, 1, 0, 2]
    v4 = [0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1]
    data = np.array([v1, v2, v3, v4]).T
    o = fwo.Oracle(data, "mi")
    lv, mv = o.levels()
    assert list(lv) == [2, 2, 3, 2] and list(mv) == [1, 1, 2, 1]
    _, lz, ctab = o.test_cond(0, 1, [2], max_k=1, want_ctab=True)
    want = np.zeros((2, 2, 3), int)
    want[0, 0, 0], want[1, 0, 0], want[0, 1, 1], want[1, 1, 1], want[1, 1, 2] = 4, 2, 2, 3, 1
    assert lz == 3
    assert sorted(map(lambda s: s.tobytes(), np.moveaxis(ctab[:2, :2, :3], 2, 0))) == \
        sorted(map(lambda s: s.tobytes(), np.moveaxis(want.astype(np.int64), 2, 0)))
    _, lz, ctab = o.test_cond(0, 1, [2, 3], max_k=2, want_ctab=True)
    want = np.zeros((2, 2, 6), int)
    want[0, 0, 0] = 2; want[0, 1, 1] = 2; want[1, 1, 1] = 2; want[0, 0, 2] = 2; want[1, 0, 2] = 2; want[1, 1, 3] = 1; want[1, 1, 4] = 1
    assert lz == 5
    got = sorted(s.tobytes() for s in np.moveaxis(ctab[:2, :2, :5], 2, 0))
    assert got == sorted(s.tobytes() for s in np.moveaxis(want[:, :, :5].astype(np.int64), 2, 0))


# ---- test/learning.jl:176-237: the 8 expected graphs ---------------------------------------
@pytest.mark.parametrize("kind", ["mi", "mi_nz", "fz", "fz_nz"])
@pytest.mark.parametrize("max_k", [0, 3])
def test_expected_graphs(kind, max_k, inputs, graphs):
    want = graphs[f"exp_{kind}_maxk{max_k}"]
    wd = {(a, b): w for a, b, w in want}
    n_obs_min = 160 if (kind.startswith("mi") and max_k == 3) else -1     # test/learning.jl:196-201
    # mode B: 1-worker single_il emulation == how the fixtures were generated (test/learning.jl:522-531)
    r = fwo.Oracle(inputs[kind], kind).lgl(max_k=max_k, n_obs_min=n_obs_min, mode="single_il")
    gd = {(a, b): w for a, b, w in r["edges"]}
    assert set(gd) == set(wd)
    for e in wd:
        assert gd[e] == pytest.approx(wd[e], rel=1e-6)
    # mode A: parallel="single" (the GPU parity target); a mode difference inside the reference
    r = fwo.Oracle(inputs[kind], kind).lgl(max_k=max_k, n_obs_min=n_obs_min, mode="single")
    ga = {(a, b) for a, b, _ in r["edges"]}
    if kind == "mi" and max_k == 3:
        assert len(ga ^ set(wd)) == 11      # absorbed by approx_nbr_diff = 22 (test/learning.jl:210-212)
    else:
        assert ga == set(wd)


def test_auto_n_obs_min(inputs):
    # learning.jl:51-61, consistent with test/learning.jl:196-201
    assert fwo.Oracle(inputs["mi"], "mi").auto_n_obs_min(3) == 160
    assert fwo.Oracle(inputs["mi"], "mi").auto_n_obs_min(0) == 20
    assert fwo.Oracle(inputs["fz"], "fz").auto_n_obs_min(3) == 20


def test_lgl_threads_deterministic(inputs):
    a = fwo.Oracle(inputs["fz"], "fz").lgl(max_k=3, mode="single", n_threads=1)
    b = fwo.Oracle(inputs["fz"], "fz").lgl(max_k=3, mode="single", n_threads=4)
    assert a["edges"] == b["edges"] and a["cond_tests"] == b["cond_tests"] == 633


def test_round5_shortcut():
    """csrc/fz.cuh evaluates round(x, digits=5) = rint(x * 1e5) / 1e5 without the divider (two fmas); the quotient must be the
    correctly rounded one for every integer the CUDA path sends through it (|k| <= 4e5), in both precisions."""
    import ctypes
    from oracle import fwo
    L = fwo.lib()
    L.fwo_round5_shortcut_mismatches.restype = ctypes.c_longlong
    L.fwo_round5_shortcut_mismatches.argtypes = [ctypes.c_longlong]
    assert L.fwo_round5_shortcut_mismatches(400000) == 0


# ---- sparse-input code path of the discrete kinds (contingency.jl:80-480) ------------------------------------------------
def _slices_equal_up_to_permutation(got, want):
    """test/contingency.jl:28-53 compare_cond_ctabs: every expected z-slice occurs among the computed ones"""
    g = [got[:2, :2, i] for i in range(got.shape[2])]
    used = set()
    for j in range(want.shape[2]):
        hit = next((i for i in range(len(g)) if i not in used and (g[i] == want[:, :, j]).all()), None)
        if hit is None:
            return False
        used.add(hit)
    return True


def test_sparse_contingency_known_answers():
    """test/contingency.jl:55-69 in its "sparse" mode: the CSC merge back-end gives the dense tables up to the order of the z slices"""
    v1 = [0, 0, 0, 0, 1, 1, 1, 1, 0, 1, 0, 1]
    v2 = [0, 0, 1, 1, 1, 1, 0, 0, 0, 1, 0, 1]
    v3 = [0, 0, 1, 1, 1, 1, 0, 0, 0, 1, 0, 2]
    v4 = [0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1]
    data = np.array([v1, v2, v3, v4]).T
    want3 = np.zeros((2, 2, 3), int)
    want3[0, 0, 0], want3[1, 0, 0], want3[0, 1, 1], want3[1, 1, 1], want3[1, 1, 2] = 4, 2, 2, 3, 1
    want34 = np.zeros((2, 2, 5), int)
    want34[0, 0, 0], want34[0, 1, 1], want34[1, 1, 1], want34[0, 0, 2], want34[1, 0, 2], want34[1, 1, 3], want34[1, 1, 4] = 2, 2, 2, 2, 2, 1, 1
    for sparse in (False, True):
        o = fwo.Oracle(data, "mi")
        o.set_sparse_semantics(sparse)
        _, lz, ctab = o.test_cond(0, 1, [2], max_k=1, want_ctab=True)
        assert lz == 3 and _slices_equal_up_to_permutation(ctab, want3)
        _, lz, ctab = o.test_cond(0, 1, [2, 3], max_k=2, want_ctab=True)
        assert lz == 5 and _slices_equal_up_to_permutation(ctab, want34)
    # test/contingency.jl:71-83: an all-zero Y under Nz must terminate
    v = np.concatenate([np.ones(25, int), np.full(25, 2)])
    A = np.stack([v, np.zeros(50, int), v], axis=1)
    o = fwo.Oracle(A, "mi_nz")
    o.set_sparse_semantics(True)
    for X, Y, Z in [(0, 1, [2]), (1, 0, [2]), (0, 2, [1]), (0, 2, [1, 1])]:
        r = o.test_cond(X, Y, Z, max_k=2)
        assert r[3] in (True, False)


def test_sparse_vs_dense_semantics(inputs):
    """test/learning.jl:369-383 ("sparse special optim (max_k 0 / 1, mi_nz)"): the mi_nz table with its last six variables made
    binary, learnt from the dense and from the sparse representation, gives the same network weights; and, test by test, the two code
    paths build the same sub-table - only levels_z of the power rule may differ (contingency.jl:171-173, 229, 461-477)."""
    A = np.array(inputs["mi_nz"], dtype=np.int32)
    A[:, -6:] = (A[:, -6:] == 0)
    for max_k in (0, 1):
        nets = []
        for sparse in (False, True):
            o = fwo.Oracle(A, "mi_nz")
            o.set_sparse_semantics(sparse)
            nets.append(o.lgl(max_k=max_k, mode="single"))
        ea, eb = {(a, b): w for a, b, w in nets[0]["edges"]}, {(a, b): w for a, b, w in nets[1]["edges"]}
        assert set(ea) == set(eb) and all(ea[e] == pytest.approx(eb[e], rel=1e-8) for e in ea)
    rng = np.random.default_rng(8)
    od, os_ = fwo.Oracle(A, "mi_nz"), fwo.Oracle(A, "mi_nz")
    os_.set_sparse_semantics(True)
    n_div = 0
    for _ in range(600):
        k = int(rng.integers(1, 4))
        v = [int(x) for x in rng.choice(A.shape[1], size=2 + k, replace=False)]
        (rd, lzd), (rs, lzs) = od.test_cond(v[0], v[1], v[2:], want_ctab=True)[:2], os_.test_cond(v[0], v[1], v[2:], want_ctab=True)[:2]
        assert lzs >= lzd                                  # the sparse path never reports fewer strata
        n_div += lzs != lzd
        if rd[3] and rs[3]:
            assert rd[0] == pytest.approx(rs[0], rel=1e-12, abs=1e-300) and rd[2] == rs[2] and rd[1] == pytest.approx(rs[1], rel=1e-10)
        else:
            assert rd[3] or not rs[3]                      # more strata can only lose power
    assert n_div > 10
