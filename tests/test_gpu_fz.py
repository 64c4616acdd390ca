"""GPU parity tests (run with -m gpu on the B200 box): the CUDA Fisher-z path, called through
the C ABI (libfwgpu.so), against the CPU oracle on the same inputs.

Bars: partial correlations (stat) bit-exact given the same Float32 cor_mat (the kernel
reproduces Julia's Float32/Float64 promotion and 5-digit rounding); p-values to 1e-12
relative (CUDA vs glibc log/erfc differ in the last ulps); decisions, subsets, test counts,
neighbour lists and edge sets exact; the tensor-core cor_mat within 1e-5 absolute of the
fp64 correlation (north_star tolerance).
"""
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


def _close_p(a, b):
    if np.isnan(a) or np.isnan(b):
        return np.isnan(a) and np.isnan(b)
    return abs(a - b) <= 1e-12 * max(abs(a), abs(b)) + 1e-300


def _same_result(g, w):
    return (g[0] == w[0] or (np.isnan(g[0]) and np.isnan(w[0]))) and _close_p(g[1], w[1]) and g[2] == w[2] and g[3] == w[3]


def _engine_with_oracle_cor(fw, x_pn):
    """Engine and oracle sharing one Float32 cor_mat (the oracle's fp64 correlation rounded to Float32)."""
    ora = fwo.Oracle(x_pn.T, "fz", cont32=True)
    cor = ora.compute_cor()
    eng = fw.Engine(0)
    eng.set_cor(cor.astype(np.float32), n_obs=x_pn.shape[1])
    return eng, ora, cor


def test_golden_conditional_tests(fw, hmp, golden_dir):
    """test(31, 21, (7,)) and test(31, 21, (7, 14, 18)) of test/tests.jl:41-74 (1-based) through the C ABI."""
    exp = json.load(open(os.path.join(golden_dir, "tests_expected.json")))
    x = np.ascontiguousarray(hmp["fz"].T)
    eng, ora, _ = _engine_with_oracle_cor(fw, x)
    got = eng.test_batch([30, 30], [20, 20], [(6,), (6, 13, 17)])
    for g, key in zip(got, ["exp_condZ1_fz", "exp_condZ3_fz"]):
        w = exp[key][0]
        assert abs(g[0] - w[0]) <= 1e-5 and abs(g[1] - w[1]) <= 5e-5 and g[2] == w[2] and g[3] == w[3]
    # univariate lookups (tests.jl:149-156) for X = 1 vs 2..50
    got = eng.test_batch([0] * 49, list(range(1, 50)))
    for g, w in zip(got, exp["exp_uni_fz"]):
        assert abs(g[0] - w[0]) <= 2e-7 and abs(g[1] - w[1]) <= 5e-7 and g[3] == w[3]
    # 1-based indices as the Julia glue passes them
    eng.L.fw_set_index_base(eng.h, 1)
    g1 = eng.test_batch([31], [21], [(7, 14, 18)])[0]
    eng.L.fw_set_index_base(eng.h, 0)
    assert g1 == eng.test_batch([30], [20], [(6, 13, 17)])[0]


@pytest.mark.parametrize("seed", [0, 1])
def test_random_conditional_tests_bit_exact(fw, synth, seed):
    rng = np.random.default_rng(seed)
    x = synth.clique(64, 300, B=8, seed=100 + seed)
    x[5] = x[4]                  # perfectly correlated pair: denominators of 0, clamps (Float64 literals in Julia)
    x[9] = -x[8]
    x[13] = 1.0                  # constant column: NaN correlations
    eng, ora, cor = _engine_with_oracle_cor(fw, x)
    X, Y, Zs = [], [], []
    for _ in range(4000):
        k = int(rng.integers(0, 4))
        v = rng.choice(64, size=2 + k, replace=False)
        X.append(int(v[0])); Y.append(int(v[1])); Zs.append(tuple(int(z) for z in v[2:]))
    # make sure the degenerate variables are exercised at every level
    for trip in [(4, 6, (5,)), (6, 7, (4, 5)), (6, 7, (4, 5, 8)), (4, 5, (6, 7, 8)), (8, 9, (1, 2)), (1, 2, (8, 9, 3)), (13, 1, (2, 3)), (1, 2, (13, 3, 4)),
                 (4, 9, (5, 8, 1)), (5, 4, (9,)), (6, 4, (5, 9, 8))]:
        X.append(trip[0]); Y.append(trip[1]); Zs.append(trip[2])
    got = eng.test_batch(X, Y, Zs, n_obs_min=20)
    n_special = 0
    for x_, y_, z_, g in zip(X, Y, Zs, got):
        w = ora.test_cond(x_, y_, list(z_), n_obs_min=20) if z_ else ora.test_uni(x_, [y_], n_obs_min=20)[0]
        assert _same_result(g, w), (x_, y_, z_, g, w)
        n_special += int(g[0] in (0.0, 1.0, -1.0) or np.isnan(g[0]))
    assert n_special >= 5
    # too few observations: TestResult(0, 1, 0, false), not an error (tests.jl:258-262)
    assert eng.test_batch([1], [2], [(3,)], n_obs_min=10_000)[0] == (0.0, 1.0, 0, False)


def test_test_subsets_matches_reference_order(fw, synth):
    rng = np.random.default_rng(7)
    x = np.concatenate([synth.clique(40, 500, B=10, seed=5), synth.chain(24, 500, B=8, seed=6)])
    eng, ora, _ = _engine_with_oracle_cor(fw, x)
    p = x.shape[0]
    jobs = []
    for _ in range(300):
        m = int(rng.integers(1, 12))
        v = rng.choice(p, size=2 + m, replace=False)
        jobs.append((int(v[0]), int(v[1]), [int(z) for z in v[2:]]))
    # within-block jobs: everything significant -> arg-max path; plus a larger one crossing the 32-slot class
    jobs += [(0, 1, list(range(2, 10))), (10, 11, list(range(12, 20))), (40, 41, [42, 43, 44]), (0, 1, list(range(2, 40)))]
    for max_k in (1, 2, 3):
        for max_tests in (0, 5, 40):
            got = eng.test_subsets_batch([j[0] for j in jobs], [j[1] for j in jobs], [j[2] for j in jobs], max_k=max_k, alpha=0.01,
                                         n_obs_min=20, max_tests=max_tests)
            for (X, Y, Z), g in zip(jobs, got):
                w = ora.test_subsets(X, Y, Z, max_k=max_k, alpha=0.01, n_obs_min=20, max_tests=max_tests)
                assert _same_result(g[0], w[0]), (X, Y, Z, max_k, max_tests, g, w)
                assert g[1] == w[1] and g[2] == w[2] and g[3] == pytest.approx(w[3], rel=1e-15), (X, Y, Z, max_k, max_tests, g, w)
    # single-job entry point + sentinels (tests.jl:285, :293-296 does not apply to fz)
    g = eng.test_subsets(0, 1, [], max_k=3)
    assert np.isnan(g[0][0]) and np.isnan(g[0][1]) and g[0][2] == -1 and g[0][3] is True and g[1] == (-1,) and g[2] == -1 and np.isnan(g[3])
    g = eng.test_subsets(0, 1, [2, 3, 4], max_k=3, n_obs_min=10_000)
    w = ora.test_subsets(0, 1, [2, 3, 4], max_k=3, n_obs_min=10_000)
    assert g[0] == w[0] == (0.0, 1.0, 0, False) and g[1] == w[1] and g[2] == w[2] == 1


def test_large_z_total_classes(fw, synth):
    """|Z_total| + 2 beyond the 64/128/224-slot shared-memory classes (global-scratch variant)."""
    x = synth.clique(300, 400, B=300, seed=11)
    eng, ora, _ = _engine_with_oracle_cor(fw, x)
    for m in (70, 140, 250):
        Z = list(range(2, 2 + m))
        g = eng.test_subsets(0, 1, Z, max_k=2, alpha=0.01, n_obs_min=20)
        w = ora.test_subsets(0, 1, Z, max_k=2, alpha=0.01, n_obs_min=20)
        assert _same_result(g[0], w[0]) and g[1] == w[1] and g[2] == w[2]


@pytest.mark.parametrize("fdr", [True, False])
def test_pairwise_stage(fw, synth, fdr):
    x = np.concatenate([synth.clique(120, 300, B=12, seed=21), synth.chain(80, 300, B=16, seed=22)])
    x[7] = 3.0                    # constant column -> NaN correlations, excluded from m (tests.jl:521-526)
    eng, ora, _ = _engine_with_oracle_cor(fw, x)
    got = eng.pw_univar_neighbors(alpha=0.01, n_obs_min=20, FDR=fdr)
    off, nbr, st, ap, rs, rp = ora.pairwise(alpha=0.01, n_obs_min=20, fdr=fdr, want_raw=True)
    assert (got.offsets == off).all() and (got.nbr == nbr).all()
    assert (got.stat == st).all()
    assert np.allclose(got.pval, ap, rtol=1e-12, atol=0)
    s = eng.pairwise_stats()
    assert s["n_tests"] == 200 * 199 // 2 and s["n_reliable"] == int((~np.isnan(rp)).sum()) and s["n_raw_sig"] == int((rp < 0.01).sum())
    # too few observations: nothing is reliable, empty lists
    got = eng.pw_univar_neighbors(alpha=0.01, n_obs_min=10_000, FDR=fdr)
    assert got.offsets[-1] == 0


def test_pairwise_golden_graph_maxk0(fw, hmp, golden_dir):
    """exp_fz_maxk0.edgelist (test/learning.jl:176-237): max_k = 0 is the FDR-filtered pairwise stage."""
    graphs = json.load(open(os.path.join(golden_dir, "learning_expected.json")))
    x = np.ascontiguousarray(hmp["fz"].T)
    eng = fw.Engine(0)
    eng.set_data_colmajor(x, "fz")
    r = eng.LGL(max_k=0)
    want = {(a, b): w for a, b, w in graphs["exp_fz_maxk0"]}
    got = {(a, b): w for a, b, w in r["edges"]}
    assert set(got) == set(want)
    for e in want:
        assert got[e] == pytest.approx(want[e], rel=1e-2)      # the reference's own tolerance (test/learning.jl:18)
        assert abs(got[e] - want[e]) <= 1e-5


def test_hiton_pc_matches_oracle(fw, synth, hmp):
    for name, x in [("hmp", np.ascontiguousarray(hmp["fz"].T)),
                    ("mix", np.concatenate([synth.clique(96, 400, B=12, seed=31), synth.chain(64, 400, B=16, seed=32)]))]:
        eng, ora, _ = _engine_with_oracle_cor(fw, x)
        p = x.shape[0]
        uni = eng.pw_univar_neighbors(alpha=0.01, n_obs_min=20)
        for max_k, max_tests in [(3, 10_000_000), (2, 10_000_000), (1, 10_000_000), (3, 25)]:
            res = eng.si_HITON_PC(np.arange(p), max_k=max_k, alpha=0.01, n_obs_min=20, max_tests=max_tests)
            tot = 0
            for T in range(p):
                a, b = uni.offsets[T], uni.offsets[T + 1]
                wn, ws, wp, wt = ora.hiton_pc(T, uni.nbr[a:b], uni.stat[a:b], uni.pval[a:b], max_k=max_k, alpha=0.01, n_obs_min=20, max_tests=max_tests)
                gn, gs, gp = res.pc(T)
                assert list(gn) == list(wn), (name, T, max_k, list(gn), list(wn))
                assert (gs == ws).all(), (name, T, max_k)
                assert np.allclose(gp, wp, rtol=1e-12, atol=0), (name, T, max_k)
                assert res.num_tests[T] == wt, (name, T, max_k, res.num_tests[T], wt)
                tot += wt
            assert res.tests_executed >= tot


@pytest.mark.parametrize("B,n", [(26, 3000), (31, 3000), (40, 6000)])
def test_hiton_pc_large_accepted_sets(fw, synth, B, n):
    """Blocks whose members all stay significant: accepted sets of B - 1 members.  B = 26 / 31 fill the 32-slot class (tables,
    p-value-free scan, one-pass colex enumeration, deferred p-values) up to its limit of 30; B = 40 starts there optimistically,
    overflows at 31 accepted members and is re-run in the 64-slot class.  Everything must equal the oracle's HITON-PC."""
    x = np.concatenate([synth.clique(2 * B, n, B=B, seed=41 + B), synth.chain(16, n, B=8, seed=43)])
    eng, ora, _ = _engine_with_oracle_cor(fw, x)
    p = x.shape[0]
    uni = eng.pw_univar_neighbors(alpha=0.01, n_obs_min=20)
    res = eng.si_HITON_PC(np.arange(p), max_k=3, alpha=0.01, n_obs_min=20)
    biggest = 0
    for T in list(range(0, 2 * B, 3)) + list(range(2 * B, p)):          # every third block member is enough (the oracle is serial)
        a, b = uni.offsets[T], uni.offsets[T + 1]
        wn, ws, wp, wt = ora.hiton_pc(T, uni.nbr[a:b], uni.stat[a:b], uni.pval[a:b], max_k=3, alpha=0.01, n_obs_min=20)
        gn, gs, gp = res.pc(T)
        assert list(gn) == list(wn), (T, list(gn), list(wn))
        assert (gs == ws).all(), T
        assert np.allclose(gp, wp, rtol=1e-12, atol=0), T
        assert res.num_tests[T] == wt, (T, res.num_tests[T], wt)
        biggest = max(biggest, len(wn))
    assert biggest >= B - 2


def test_lgl_golden_graph_maxk3(fw, hmp, golden_dir):
    """exp_fz_maxk3.edgelist: parallel="single" recovers the identical edge set."""
    graphs = json.load(open(os.path.join(golden_dir, "learning_expected.json")))
    x = np.ascontiguousarray(hmp["fz"].T)
    eng = fw.Engine(0)
    eng.set_data_colmajor(x, "fz")
    r = eng.LGL(max_k=3)
    want = {(a, b) for a, b, _ in graphs["exp_fz_maxk3"]}
    assert {(a, b) for a, b, _ in r["edges"]} == want
    # and the whole run equals the oracle's "single" mode on the engine's own cor_mat
    ora = fwo.Oracle(x.T, "fz", cont32=True)
    ora.set_cor(eng.cor().astype(np.float64))
    w = ora.lgl(max_k=3, mode="single")
    assert [(a, b) for a, b, _ in r["edges"]] == [(a, b) for a, b, _ in w["edges"]]
    assert max(abs(g[2] - ww[2]) for g, ww in zip(r["edges"], w["edges"])) == 0.0
    assert r["cond_tests"] == w["cond_tests"]


def test_cor_matrix_tolerance(fw, synth):
    for (p, n, seed) in [(50, 346, 1), (333, 1000, 2), (1024, 2000, 3)]:
        x = synth.clique(p, n, B=16, seed=seed)
        x[3] *= 1000.0
        x[4] += 50.0              # large mean: the two-pass standardisation must not cancel
        eng = fw.Engine(0)
        eng.set_data_colmajor(x, "fz")
        got = eng.cor()
        want = np.corrcoef(x.astype(np.float64))
        assert np.abs(got - want).max() <= 1e-5          # north_star tolerance on correlations
        assert (np.diag(got) == 1.0).all() and (got == got.T).all() and np.abs(got).max() <= 1.0
        # upload hidden behind the GEMM (column bands): same matrix bit for bit
        eng2 = fw.Engine(0)
        got2 = eng2.upload_and_cor(x, want_host=True)
        assert (got2 == got).all()
        assert eng2.test_batch([0], [1], [(2, 3)])[0] == eng.test_batch([0], [1], [(2, 3)])[0]
    # constant column -> NaN row/column, unit diagonal (Statistics.cov2cor!)
    x = synth.clique(40, 200, B=8, seed=4)
    x[5] = 2.0
    eng = fw.Engine(0)
    eng.set_data_colmajor(x, "fz")
    got = eng.cor()
    assert np.isnan(got[5, :5]).all() and np.isnan(got[6:, 5]).all() and got[5, 5] == 1.0


def test_end_to_end_edge_recovery(fw, synth):
    """Chain workload: GPU pipeline == oracle pipeline on the engine's cor_mat, and the
    true skeleton (chain edges) is recovered."""
    x = synth.chain(512, 2000, B=32, seed=synth.BASE_SEED + 1)
    eng = fw.Engine(0)
    eng.set_data_colmajor(x, "fz")
    r = eng.LGL(max_k=3)
    ora = fwo.Oracle(x.T, "fz", cont32=True)
    ora.set_cor(eng.cor().astype(np.float64))
    w = ora.lgl(max_k=3, mode="single", n_threads=8)
    assert [(a, b) for a, b, _ in r["edges"]] == [(a, b) for a, b, _ in w["edges"]]
    assert r["cond_tests"] == w["cond_tests"]
    truth = {(v - 1, v) for v in range(512) if v % 32 != 0}
    got = {(a, b) for a, b, _ in r["edges"]}
    assert len(truth - got) == 0 and len(got - truth) <= len(truth) // 10


def test_tensor_core_cor_mat_vs_fp64_derived(fw, synth, hmp):
    """The engine's cor_mat comes from a split-bf16 tensor-core GEMM (<= 2.4e-6 from fp64); the reference's is Float32(cor in
    Float32/fp64) (3e-8).  pcor_rec rounds numerators to 1e-5, so single roundings can flip against the reference.  This
    bounds the effect end to end: the oracle on the fp64-derived Float32 cor_mat versus the engine on its own matrix - edge sets
    differ in at most a borderline edge or two per thousand, common edges carry weights within 5e-5."""
    tables = [np.ascontiguousarray(hmp["fz"].T),
              np.concatenate([synth.clique(360, 700, B=12, seed=31), synth.chain(240, 700, B=20, seed=32)])]
    for x in tables:
        eng = fw.Engine(0)
        eng.set_data_colmajor(x, "fz")
        r = eng.LGL(max_k=3)
        ora = fwo.Oracle(x.T, "fz", cont32=True)
        c64 = ora.compute_cor()                                     # Float32(cor in fp64)
        assert np.abs(eng.cor() - c64).max() <= 3e-6
        w = ora.lgl(max_k=3, mode="single", n_threads=8)
        ge = {(a, b): v for a, b, v in r["edges"]}
        we = {(a, b): v for a, b, v in w["edges"]}
        assert len(set(ge) ^ set(we)) <= max(1, len(we) // 250), (len(set(ge) ^ set(we)), len(we))
        common = set(ge) & set(we)
        assert max(abs(ge[e] - we[e]) for e in common) <= 5e-5
        assert abs(r["cond_tests"] - w["cond_tests"]) <= max(50, w["cond_tests"] // 100)
