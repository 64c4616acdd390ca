"""Size-independent properties at BASELINE.json's full size (50 000 x 10 000, fz, max_k = 3), plus tiny / degenerate inputs and
the error behaviour of the C ABI (no silent fallbacks)."""
import numpy as np
import pytest

import fwload

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fw():
    return fwload.load()


@pytest.fixture(scope="module")
def synth():
    return fwload.load_sub("synth")


def test_full_size_properties(fw, synth):
    p, n, B = 50000, 10000, 24
    x = synth.clique(p, n, B=B, seed=synth.BASE_SEED + 3)
    eng = fw.Engine(0)
    eng.set_data_colmajor(x, "fz")
    eng.cor(want_host=False)
    uni = eng.pw_univar_neighbors(alpha=0.01, n_obs_min=20)
    st = eng.pairwise_stats()
    assert st["n_tests"] == p * (p - 1) // 2 == st["n_reliable"]
    # neighbour lists: ascending, symmetric, every within-block pair present (r = 0.64 at n = 10 000 is overwhelming)
    off, nbr = uni.offsets, uni.nbr
    blk = np.minimum(B, p - (np.arange(p) // B) * B)                      # block size of every variable (the last block is short)
    assert (np.diff(off) >= blk - 1).all()
    rows = np.repeat(np.arange(p), np.diff(off))
    assert (rows != nbr).all()
    key = rows * p + nbr
    assert (np.diff(key) > 0).all()                                     # sorted by (row, neighbour): ascending neighbour index
    rev = nbr * p + rows
    pos = np.searchsorted(key, rev)
    assert (key[pos] == rev).all()                                       # (a, b) present iff (b, a)
    assert (uni.stat[pos] == uni.stat).all() and (uni.pval[pos] == uni.pval).all()
    same_block = (rows // B) == (nbr // B)
    assert same_block.sum() == int((blk - 1).sum())                       # every within-block pair is a univariate neighbour
    assert (np.abs(uni.stat) <= 1.0).all() and (uni.pval < 0.01).all()
    # HITON-PC on 4 000 random targets, in two different batchings: identical results (sharding invariance / idempotence)
    rng = np.random.default_rng(0)
    tg = np.sort(rng.choice(p, size=4000, replace=False))
    r1 = eng.si_HITON_PC(tg, max_k=3, alpha=0.01, n_obs_min=20, want_tpc=False)
    perm = rng.permutation(len(tg))
    r2 = eng.si_HITON_PC(tg[perm][:2500], max_k=3, alpha=0.01, n_obs_min=20, want_tpc=False)
    n_pc = n_out = 0
    for j in range(2500):
        i = int(perm[j])
        a1, s1, p1 = r1.pc(i)
        a2, s2, p2 = r2.pc(j)
        assert (a1 == a2).all() and (s1 == s2).all() and (p1 == p2).all() and r1.num_tests[i] == r2.num_tests[j]
    for i, T in enumerate(tg):
        nb, s, pv = r1.pc(i)
        cand = set(uni.nbr[off[T]:off[T + 1]].tolist())
        assert set(nb.tolist()) <= cand                                  # PC is a subset of the univariate neighbours
        inb = (nb // B) == (T // B)
        assert inb.sum() == blk[T] - 1                                   # the true skeleton: every block-mate stays (partial r ~ 0.22)
        n_out += int((~inb).sum())                                       # FDR-surviving false positives may survive conditioning too
        assert (pv < 0.01).all() and (np.abs(s[inb]) > 0.05).all()
        n_pc += len(nb)
    assert n_out <= 0.01 * n_pc                                          # at most the nominal FDR level (observed: 0.2 %)
    # test counts: interleaving over B-1 candidates + elimination, every subset evaluated (nothing exits early)
    from math import comb
    m = B - 1
    per_target = comb(m, 4) + comb(m, 3) + comb(m, 2) + m * (comb(m - 1, 3) + comb(m - 1, 2) + (m - 1))
    clean = (np.diff(off)[tg] == B - 1) & (blk[tg] == B)                  # full blocks without false-positive univariate neighbours
    assert (r1.num_tests[clean] == per_target).all() and clean.sum() > 3000


def test_tiny_and_degenerate_inputs(fw):
    rng = np.random.default_rng(1)
    # p = 2, n = 5
    x = rng.standard_normal((2, 5)).astype(np.float32)
    eng = fw.Engine(0)
    eng.set_data_colmajor(x, "fz")
    c = eng.cor()
    assert c.shape == (2, 2) and c[0, 0] == 1 and abs(c[0, 1] - np.corrcoef(x.astype(np.float64))[0, 1]) < 1e-5
    r = eng.LGL(max_k=3, n_obs_min=0)
    assert r["cond_tests"] == 0 and len(r["edges"]) <= 1
    # n = 3: the Fisher-z sample factor n - 3 is 0 -> z = 0 -> p = 1 for every test (statfuns.jl:6-10)
    x = rng.standard_normal((6, 3)).astype(np.float32)
    eng = fw.Engine(0)
    eng.set_data_colmajor(x, "fz")
    eng.cor(want_host=False)
    assert all(t[1] == 1.0 for t in eng.test_batch([0, 1], [1, 2], [(2,), (0, 3)]))
    assert eng.pw_univar_neighbors(alpha=0.01, n_obs_min=0).offsets[-1] == 0
    # a table where nothing is associated: no candidates, HITON-PC returns empty lists
    x = rng.standard_normal((40, 200)).astype(np.float32)
    eng = fw.Engine(0)
    eng.set_data_colmajor(x, "fz")
    r = eng.LGL(max_k=3)
    assert r["cond_tests"] == 0 and r["hiton"].pc_count.sum() == 0


def test_abi_errors_are_loud(fw):
    eng = fw.Engine(0)
    with pytest.raises(fw.FwError, match="no cor_mat"):
        eng.test_batch([0], [1], [(2,)], kind="fz")
    with pytest.raises(fw.FwError, match="no discrete table"):
        eng.test_batch([0], [1], [(2,)], kind="mi")
    x = np.random.default_rng(2).standard_normal((10, 50)).astype(np.float32)
    eng.set_data_colmajor(x, "fz")
    eng.cor(want_host=False)
    with pytest.raises(fw.FwError, match="out of range"):
        eng.test_batch([0], [10], [(2,)])
    with pytest.raises(fw.FwError, match="max_k"):
        eng.test_subsets(0, 1, [2, 3, 4, 5, 6], max_k=4)
    with pytest.raises(fw.FwError, match="neighbour lists"):
        eng.si_HITON_PC([0, 1])
    # a new table (same p) invalidates the neighbour lists and the cor_mat of the previous one: loud, not stale
    eng.pw_univar_neighbors(alpha=0.5, n_obs_min=0)
    eng.si_HITON_PC([0, 1], max_k=1)
    eng.set_data_colmajor(x[::-1].copy(), "fz")
    with pytest.raises(fw.FwError, match="no cor_mat"):
        eng.test_batch([0], [1], [(2,)])
    eng.cor(want_host=False)
    with pytest.raises(fw.FwError, match="neighbour lists"):
        eng.si_HITON_PC([0, 1], max_k=1)
    with pytest.raises(fw.FwError, match="neighbour lists"):
        eng.univar_nbrs()
    xi = (np.random.default_rng(3).random((10, 50)) < 0.5).astype(np.int32)
    eng.set_data_colmajor(xi, "mi")
    eng.pw_univar_neighbors(alpha=0.5, n_obs_min=0)
    eng.set_data_colmajor(xi[::-1].copy(), "mi")
    with pytest.raises(fw.FwError, match="neighbour lists"):
        eng.si_HITON_PC([0, 1], max_k=1)
    with pytest.raises(fw.FwError, match="levels"):
        e2 = fw.Engine(0)
        e2.set_data_colmajor(np.arange(60).reshape(3, 20).astype(np.int32) % 7, "mi")      # 7 levels > 4
    with pytest.raises(fw.FwError):
        fw.Engine(99)
