"""Whitelists, blacklists and rejection records of si_HITON_PC through the C ABI (fw_hiton_pc_ex; src/hiton.jl:20-38, 72-74),
driven by the reference's default schedule: the host runs the 1-worker single_il loop of src/interleaved.jl (feed-forward
whitelists) with every target job executed on the GPU, and must reproduce all 8 committed graphs of
test/data/learning_expected - edges AND every weight - which were generated in exactly that mode (test/learning.jl:522-531)."""
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
def hmp(golden_dir):
    return np.load(os.path.join(golden_dir, "hmp_inputs.npz"))


@pytest.fixture(scope="module")
def graphs(golden_dir):
    return json.load(open(os.path.join(golden_dir, "learning_expected.json")))


def _engine(fw, hmp, kind):
    x = np.ascontiguousarray(hmp[kind].T)
    eng = fw.Engine(0)
    eng.set_data_colmajor(x.astype(np.float32 if kind.startswith("fz") else np.int32), kind)
    ora = fwo.Oracle(hmp[kind], kind)
    if kind == "fz":
        eng.set_cor(ora.compute_cor())            # the reference's Float32(cor in fp64), so that weights can be compared digit by digit
    return eng, ora


@pytest.mark.parametrize("kind", ["mi", "mi_nz", "fz", "fz_nz"])
@pytest.mark.parametrize("max_k", [0, 3])
def test_single_il_schedule_reproduces_the_committed_graphs(fw, hmp, graphs, kind, max_k, tmp_path, golden_dir):
    eng, ora = _engine(fw, hmp, kind)
    nom = 160 if (kind.startswith("mi") and max_k == 3) else -1                  # test/learning.jl:196-201
    r = eng.LGL(max_k=max_k, n_obs_min=nom, parallel="single_il")
    want = {(a, b): w for a, b, w in graphs[f"exp_{kind}_maxk{max_k}"]}
    got = {(a, b): w for a, b, w in r["edges"]}
    assert set(got) == set(want)                                                 # incl. mi / max_k = 3: the 81 edges of the fixture
    for e in want:
        assert got[e] == pytest.approx(want[e], rel=1e-6), e
    # and the oracle's own emulation of the same schedule: identical edges, weights and test count
    w = ora.lgl(max_k=max_k, n_obs_min=nom, mode="single_il")
    assert [(a, b) for a, b, _ in r["edges"]] == [(a, b) for a, b, _ in w["edges"]]
    tol = 0.0 if kind.startswith("fz") else 1e-12
    assert np.allclose([e[2] for e in r["edges"]], [e[2] for e in w["edges"]], rtol=tol, atol=0.0)
    assert r["cond_tests"] == w["cond_tests"]
    # written out: same lines as the reference's file (names and order identical, weights to 1e-6)
    out = tmp_path / "net.edgelist"
    fw.write_edgelist(str(out), r["edges"], p=50)
    ref_lines = open(os.path.join(golden_dir, "edgelists", f"exp_{kind}_maxk{max_k}.edgelist")).read().split("\n")
    my_lines = open(out).read().split("\n")
    assert my_lines[:2] == ref_lines[:2] and len(my_lines) == len(ref_lines)
    for a, b in zip(my_lines[2:], ref_lines[2:]):
        if b:
            assert a.split("\t")[:2] == b.split("\t")[:2] and float(a.split("\t")[2]) == pytest.approx(float(b.split("\t")[2]), rel=1e-6)


@pytest.mark.parametrize("kind", ["mi", "fz", "fz_nz"])
def test_whitelist_blacklist_against_oracle(fw, hmp, kind):
    """arbitrary per-target whitelists (incl. the first candidate, variables that are no candidates at all) and blacklists"""
    eng, ora = _engine(fw, hmp, kind)
    nom = 160 if kind.startswith("mi") else 20
    uni = eng.pw_univar_neighbors(alpha=0.01, n_obs_min=nom)
    rng = np.random.default_rng(4)
    targets = [int(t) for t in np.nonzero(uni.degree() >= 3)[0]][:30]
    wls, bls = [], []
    for T in targets:
        nb = uni.nbr[uni.offsets[T]:uni.offsets[T + 1]]
        k = int(rng.integers(0, min(4, len(nb)) + 1))
        wl = set(int(v) for v in rng.choice(nb, size=k, replace=False)) | {int(rng.integers(0, 50))}
        wls.append(sorted(wl - {T}))
        bls.append([])
    res = eng.si_HITON_PC(targets, max_k=3, alpha=0.01, n_obs_min=nom, whitelists=wls, blacklists=bls)
    tol = 0.0 if kind.startswith("fz") else 1e-12
    for i, T in enumerate(targets):
        a, b = uni.offsets[T], uni.offsets[T + 1]
        wn, ws, wp, wt = ora.hiton_pc(T, uni.nbr[a:b], uni.stat[a:b], uni.pval[a:b], max_k=3, alpha=0.01, n_obs_min=nom, whitelist=wls[i])
        gn, gs, gp = res.pc(i)
        assert list(gn) == list(wn) and res.num_tests[i] == wt, (T, wls[i], list(gn), list(wn))
        assert np.allclose(gs, ws, rtol=tol, atol=0.0, equal_nan=True) and np.allclose(gp, wp, rtol=1e-9, atol=1e-300, equal_nan=True)
    # a blacklisted candidate is never tested and never accepted; everything else behaves as if it were not a candidate
    T = targets[0]
    nb = uni.nbr[uni.offsets[T]:uni.offsets[T + 1]]
    base = eng.si_HITON_PC([T], max_k=3, alpha=0.01, n_obs_min=nom)
    if base.pc_count[0] >= 2:
        drop = int(base.pc(0)[0][0])
        r_bl = eng.si_HITON_PC([T], max_k=3, alpha=0.01, n_obs_min=nom, blacklists=[[drop]])
        keep = uni.nbr[uni.offsets[T]:uni.offsets[T + 1]] != drop
        eng2, _ = _engine(fw, hmp, kind)
        off2 = uni.offsets.copy(); off2[T + 1:] -= 1
        sel = np.ones(len(uni.nbr), bool); sel[uni.offsets[T]:uni.offsets[T + 1]] = keep
        eng2.set_univar_nbrs(off2, uni.nbr[sel], uni.stat[sel], uni.pval[sel])
        r_wo = eng2.si_HITON_PC([T], max_k=3, alpha=0.01, n_obs_min=nom)
        assert drop not in r_bl.pc(0)[0]
        assert list(r_bl.pc(0)[0]) == list(r_wo.pc(0)[0]) and r_bl.num_tests[0] == r_wo.num_tests[0]
        assert np.array_equal(r_bl.pc(0)[1], r_wo.pc(0)[1])


@pytest.mark.parametrize("kind", ["mi", "fz", "fz_nz"])
def test_rejection_records(fw, hmp, kind):
    """track_rejections: every candidate is accepted or rejected; a rejection record holds the subset and the TestResult that
    rejected the candidate (hiton.jl:72-74), reproducible as a single test, and the same (num_tests, frac) test_subsets reports."""
    eng, ora = _engine(fw, hmp, kind)
    nom = 160 if kind.startswith("mi") else 20
    uni = eng.pw_univar_neighbors(alpha=0.01, n_obs_min=nom)
    targets = [int(t) for t in np.nonzero(uni.degree() >= 2)[0]][:25]
    res = eng.si_HITON_PC(targets, max_k=3, alpha=0.01, n_obs_min=nom, track_rejections=True)
    plain = eng.si_HITON_PC(targets, max_k=3, alpha=0.01, n_obs_min=nom)
    n_rej = 0
    for i, T in enumerate(targets):
        assert list(res.pc(i)[0]) == list(plain.pc(i)[0]) and np.array_equal(res.pc(i)[1], plain.pc(i)[1]) and res.num_tests[i] == plain.num_tests[i]
        rej = res.rejections(i)
        cands = set(int(v) for v in uni.nbr[uni.offsets[T]:uni.offsets[T + 1]])
        assert set(rej) | set(int(v) for v in res.pc(i)[0]) == cands and not (set(rej) & set(int(v) for v in res.pc(i)[0]))
        for cand, (Zs, tr, (nt, frac)) in rej.items():
            n_rej += 1
            assert not (tr[1] < 0.01 and tr[3]) and nt >= 1 and 0.0 < frac <= 1.0 and 1 <= len(Zs) <= 3
            if kind == "fz_nz":
                continue                           # the sub-correlations depend on the whole accepted list of that call (cor_subset!)
            single = eng.test(T, cand, Zs, n_obs_min=nom)
            assert single[2] == tr[2] and single[3] == tr[3]
            assert abs(single[0]) == pytest.approx(abs(tr[0]), rel=1e-12, abs=1e-300) and single[1] == pytest.approx(tr[1], rel=1e-9, abs=1e-300)
    assert n_rej > (10 if kind != "fz_nz" else 0)        # the fz_nz fixture has 10 univariate edges in all
