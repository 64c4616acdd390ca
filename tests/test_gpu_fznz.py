"""GPU parity tests of the zero-ignoring Fisher-z path (fz_nz) through the C ABI against the CPU oracle.
The per-job sub-correlations follow ONE canonical summation order on both sides (oracle/fw_oracle.cpp::cor_view,
csrc/fznz.cuh), so the Float32 correlations are bit-equal and the bars are those of the plain Fisher-z path: statistic
bit-exact, p-value to 1e-12 relative (CUDA vs glibc log/erfc), identical decisions, subsets and test counts."""
import json
import os

import numpy as np
import pytest

import fwload
from oracle import fwo

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fw():
    return fwload.load()


@pytest.fixture(scope="module")
def synth():
    return fwload.load_sub("synth")


@pytest.fixture(scope="module")
def hmp(golden_dir):
    return np.load(os.path.join(golden_dir, "hmp_inputs.npz"))


def _cmp(g, w, stats):
    assert g[2] == w[2] and g[3] == w[3], (g, w)
    if np.isnan(w[0]):
        assert np.isnan(g[0]) and np.isnan(g[1])
        return
    assert g[0] == w[0], (g, w)                                   # bit-exact statistic
    stats["exact"] += 1
    assert abs(g[1] - w[1]) <= 1e-12 * max(abs(w[1]), 1e-300) + 1e-300, (g, w)
    stats["n"] += 1


def _table(synth, seed, n=500):
    lat = np.concatenate([synth.clique(40, n, B=8, seed=seed), synth.chain(24, n, B=8, seed=seed + 1)])
    x = synth.with_zeros(lat, zero_frac=0.35, seed=seed)
    x[3] = 0.0                                     # never present
    x[5, : n - 10] = 0.0                           # present in 10 samples only (< n_obs_min)
    x[7] = np.where(x[7] != 0, 1.5, 0.0)           # constant where present: NaN correlations
    return x


def test_golden_fznz(fw, hmp, golden_dir):
    exp = json.load(open(os.path.join(golden_dir, "tests_expected.json")))
    x = np.ascontiguousarray(hmp["fz_nz"].T)
    eng = fw.Engine(0)
    eng.set_data_colmajor(x, "fz_nz")
    got = eng.test_batch([0] * 49, list(range(1, 50)))
    for g, w in zip(got, exp["exp_uni_fz_nz"]):
        assert abs(g[0] - w[0]) <= 2e-7 and abs(g[1] - w[1]) <= 5e-7 and g[3] == w[3], (g, w)
    got = eng.test_batch([30, 30], [20, 20], [(6,), (6, 13, 17)])
    for g, key in zip(got, ["exp_condZ1_fz_nz", "exp_condZ3_fz_nz"]):
        w = exp[key][0]
        assert abs(g[0] - w[0]) <= 1e-5 and abs(g[1] - w[1]) <= 5e-5 and g[2] == w[2] and g[3] == w[3], (g, w)


def test_random_fznz_tests(fw, synth):
    x = _table(synth, 80)
    eng = fw.Engine(0)
    eng.set_data_colmajor(x, "fz_nz")
    ora = fwo.Oracle(x.T, "fz_nz")
    rng = np.random.default_rng(5)
    p = x.shape[0]
    X, Y, Zs = [], [], []
    for _ in range(1200):
        k = int(rng.integers(0, 4))
        v = rng.choice(p, size=2 + k, replace=False)
        X.append(int(v[0])); Y.append(int(v[1])); Zs.append(tuple(int(z) for z in v[2:]))
    for trip in [(3, 1, ()), (1, 3, ()), (5, 1, ()), (1, 5, (2,)), (7, 1, ()), (1, 7, (2, 4)), (1, 2, (7, 4, 6)), (3, 5, (1,))]:
        X.append(trip[0]); Y.append(trip[1]); Zs.append(trip[2])
    for nom in (20, 0):
        got = eng.test_batch(X, Y, Zs, n_obs_min=nom)
        stats = {"n": 0, "exact": 0}
        for x_, y_, z_, g in zip(X, Y, Zs, got):
            w = ora.test_cond(x_, y_, list(z_), n_obs_min=nom) if z_ else ora.test_uni(x_, [y_], n_obs_min=nom)[0]
            _cmp(g, w, stats)
        assert stats["exact"] == stats["n"]


def test_fznz_pairwise_subsets_hiton(fw, synth):
    x = _table(synth, 90)
    eng = fw.Engine(0)
    eng.set_data_colmajor(x, "fz_nz")
    ora = fwo.Oracle(x.T, "fz_nz")
    p = x.shape[0]
    got = eng.pw_univar_neighbors(alpha=0.01, n_obs_min=20)
    off, nbr, st, ap, rs, rp = ora.pairwise(alpha=0.01, n_obs_min=20, want_raw=True)
    assert (got.offsets == off).all() and (got.nbr == nbr).all()
    assert (got.stat == st).all() and np.allclose(got.pval, ap, rtol=1e-12, atol=1e-300)
    s = eng.pairwise_stats()
    assert s["n_reliable"] == int((~np.isnan(rp)).sum()) and s["n_raw_sig"] == int((rp < 0.01).sum())
    # subset search
    rng = np.random.default_rng(2)
    jobs = []
    for _ in range(120):
        m = int(rng.integers(1, 8))
        v = rng.choice(p, size=2 + m, replace=False)
        jobs.append((int(v[0]), int(v[1]), [int(z) for z in v[2:]]))
    jobs += [(0, 1, [2, 4, 6]), (5, 1, [2, 4]), (8, 9, [10, 11, 12, 13, 14, 15])]
    for max_k, nom in [(3, 20), (2, 20), (3, 150)]:
        gres = eng.test_subsets_batch([j[0] for j in jobs], [j[1] for j in jobs], [j[2] for j in jobs], max_k=max_k, alpha=0.01, n_obs_min=nom)
        stats = {"n": 0, "exact": 0}
        for (X, Y, Z), g in zip(jobs, gres):
            w = ora.test_subsets(X, Y, Z, max_k=max_k, alpha=0.01, n_obs_min=nom)
            _cmp(g[0], w[0], stats)
            assert g[1] == w[1] and g[2] == w[2], (X, Y, Z, g, w)          # same subset, same num_tests
    # HITON-PC per target
    uni = eng.univar_nbrs()
    res = eng.si_HITON_PC(np.arange(p), max_k=3, alpha=0.01, n_obs_min=20)
    n_same = 0
    for T in range(p):
        a, b = uni.offsets[T], uni.offsets[T + 1]
        wn, ws, wp, wt = ora.hiton_pc(T, uni.nbr[a:b], uni.stat[a:b], uni.pval[a:b], max_k=3, alpha=0.01, n_obs_min=20)
        gn, gs, gp = res.pc(T)
        if list(gn) == list(wn) and res.num_tests[T] == wt:
            n_same += 1
            assert (np.asarray(gs) == np.asarray(ws)).all() and np.allclose(gp, wp, rtol=1e-12, atol=1e-300)
    assert n_same == p


def test_fznz_subsets_gram_sizes(fw, synth):
    """Jobs of 3 ... 34 variables: the register-blocked Gram takes one pass up to 27 variables, two passes for 28 ... 32, and
    the pair-per-warp fallback beyond (64-slot class)."""
    lat = synth.clique(48, 600, B=48, seed=61)
    x = synth.with_zeros(lat, zero_frac=0.2, seed=62)
    eng = fw.Engine(0)
    eng.set_data_colmajor(x, "fz_nz")
    ora = fwo.Oracle(x.T, "fz_nz")
    jobs = [(0, 1, list(range(2, 2 + m))) for m in (1, 2, 5, 12, 24, 25, 26, 27, 28, 29, 30, 32, 40)]
    gres = eng.test_subsets_batch([j[0] for j in jobs], [j[1] for j in jobs], [j[2] for j in jobs], max_k=2, alpha=0.01, n_obs_min=20)
    stats = {"n": 0, "exact": 0}
    for (X, Y, Z), g in zip(jobs, gres):
        w = ora.test_subsets(X, Y, Z, max_k=2, alpha=0.01, n_obs_min=20)
        _cmp(g[0], w[0], stats)
        assert g[2] == w[2], (len(Z), g, w)            # num_tests
    assert stats["exact"] == stats["n"]


def test_fznz_pairwise_prefilter_equals_exhaustive(fw, synth):
    """The tensor-core pre-filter + exact candidate test must give the same neighbour lists, reliable-test count and
    raw-significant count as the exact test on every pair (FWGPU_FZNZ_TC=0), on a heterogeneous table with meta variables."""
    x, _ = synth.hetero(1210, 3000, B=24, seed=11)
    x[17] = np.where(x[17] != 0, 2.5, 0.0)          # constant where present
    x[40, 30:] = 0.0                                # present in < n_obs_min samples
    res = {}
    for mode in ("1", "0"):
        os.environ["FWGPU_FZNZ_TC"] = mode
        try:
            eng = fw.Engine(0)
            eng.set_data_colmajor(x, "fz_nz")
            for alpha, nom in ((0.01, 20), (0.05, 100)):
                got = eng.pw_univar_neighbors(alpha=alpha, n_obs_min=nom)
                res[(mode, alpha)] = (got.offsets.copy(), got.nbr.copy(), got.stat.copy(), got.pval.copy(), dict(eng.pairwise_stats()))
        finally:
            os.environ.pop("FWGPU_FZNZ_TC", None)
    for alpha in (0.01, 0.05):
        a, b = res[("1", alpha)], res[("0", alpha)]
        assert (a[0] == b[0]).all() and (a[1] == b[1]).all()
        assert (a[2] == b[2]).all() and (a[3] == b[3]).all()          # both paths take the statistic from the same exact kernel
        assert a[4] == b[4], (a[4], b[4])
        assert a[4]["n_raw_sig"] > 1000


def test_fznz_pairwise_prefilter_adversarial(fw):
    """Views far from the column mean with a tiny variance, heavy tails and few shared rows: the bf16 moments of the pre-filter
    cancel there, so only its error intervals keep truly significant pairs.  Must equal the exhaustive path."""
    rng = np.random.default_rng(77)
    n, p = 2000, 384
    x = np.zeros((p, n), np.float32)
    half = n // 2
    for v in range(p):
        kind = v % 4
        f = rng.standard_normal(n)
        if kind == 0:
            # bimodal: +50 on the first half, -50 on the second; block-mates share a small signal inside each mode
            base = np.where(np.arange(n) < half, 50.0, -50.0)
            g = rng.standard_normal(n) if v % 8 else np.zeros(n)
            col = base + 0.05 * (0.9 * np.roll(f, 0) + 0.4 * g)
            present = rng.random(n) < 0.6
            if v % 8 == 0:
                present &= np.arange(n) < half            # only ever seen in the +50 mode
        elif kind == 1:
            col = np.exp(3.0 * f)                          # heavy tail
            present = rng.random(n) < 0.5
        elif kind == 2:
            col = 1000.0 + 0.01 * f                        # huge offset, tiny variance
            present = rng.random(n) < 0.15
        else:
            col = f
            present = rng.random(n) < 0.03                 # few rows: views near n_obs_min
        x[v] = np.where(present, col, 0.0).astype(np.float32)
    # correlated partners inside the difficult regimes
    shared = rng.standard_normal(n)
    for a, b in ((0, 8), (2, 6), (1, 5), (16, 24), (10, 14)):
        for v in (a, b):
            nzv = x[v] != 0
            x[v] = np.where(nzv, x[v] + (0.04 if v % 4 == 0 else (0.008 if v % 4 == 2 else 0.5 * np.abs(x[v]))) * shared, 0.0).astype(np.float32)
    res = {}
    for mode in ("1", "0"):
        os.environ["FWGPU_FZNZ_TC"] = mode
        try:
            eng = fw.Engine(0)
            eng.set_data_colmajor(x, "fz_nz")
            for alpha, nom in ((0.01, 20), (0.05, 5)):
                got = eng.pw_univar_neighbors(alpha=alpha, n_obs_min=nom)
                res[(mode, alpha)] = (got.offsets.copy(), got.nbr.copy(), got.stat.copy(), got.pval.copy(), dict(eng.pairwise_stats()))
        finally:
            os.environ.pop("FWGPU_FZNZ_TC", None)
    for alpha in (0.01, 0.05):
        a, b = res[("1", alpha)], res[("0", alpha)]
        assert a[4] == b[4], (a[4], b[4])
        assert (a[0] == b[0]).all() and (a[1] == b[1]).all() and (a[2] == b[2]).all() and (a[3] == b[3]).all()
        assert a[4]["n_raw_sig"] > 50
    # and the exhaustive path itself against the oracle
    ora = fwo.Oracle(x.T, "fz_nz")
    off, nbr, st, ap = ora.pairwise(alpha=0.01, n_obs_min=20)
    a = res[("0", 0.01)]
    assert (a[0] == off).all() and (a[1] == nbr).all() and (a[2] == st).all() and np.allclose(a[3], ap, rtol=1e-12, atol=1e-300)


def test_fznz_pairwise_prefilter_candidate_overflow(fw, synth):
    """One big block: nearly every pair is significant, so the candidate list outgrows its first allocation (n_pairs / 16) and
    the pre-filter is re-run with the exact size; the result must still equal the exhaustive path."""
    lat = synth.clique(1600, 800, B=1600, seed=71)
    x = synth.with_zeros(lat, zero_frac=0.3, seed=72)
    res = {}
    for mode in ("1", "0"):
        os.environ["FWGPU_FZNZ_TC"] = mode
        try:
            eng = fw.Engine(0)
            eng.set_data_colmajor(x, "fz_nz")
            got = eng.pw_univar_neighbors(alpha=0.01, n_obs_min=20)
            res[mode] = (got.offsets.copy(), got.nbr.copy(), got.stat.copy(), got.pval.copy(), dict(eng.pairwise_stats()))
        finally:
            os.environ.pop("FWGPU_FZNZ_TC", None)
    a, b = res["1"], res["0"]
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all() and (a[2] == b[2]).all() and (a[3] == b[3]).all() and a[4] == b[4]
    assert a[4]["n_raw_sig"] > 1600 * 1599 // 2 // 2          # far more candidates than n_pairs / 16 + 1024


def test_fznz_golden_graphs(fw, hmp, golden_dir):
    graphs = json.load(open(os.path.join(golden_dir, "learning_expected.json")))
    x = np.ascontiguousarray(hmp["fz_nz"].T)
    eng = fw.Engine(0)
    eng.set_data_colmajor(x, "fz_nz")
    for max_k in (0, 3):
        r = eng.LGL(max_k=max_k)
        want = {(a, b): w for a, b, w in graphs[f"exp_fz_nz_maxk{max_k}"]}
        got = {(a, b): w for a, b, w in r["edges"]}
        assert set(got) == set(want)                     # "single" mode recovers the identical edge set
    w3 = fwo.Oracle(x.T, "fz_nz").lgl(max_k=3, mode="single")
    assert [(a, b) for a, b, _ in r["edges"]] == [(a, b) for a, b, _ in w3["edges"]]
    assert [e[2] for e in r["edges"]] == [e[2] for e in w3["edges"]]          # weights bit-equal
    assert r["cond_tests"] == w3["cond_tests"] == 57
