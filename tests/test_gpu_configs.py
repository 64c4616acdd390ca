"""GPU-vs-oracle parity at the BASELINE.json configuration shapes (C2, C3, a C5-shaped table; C4 is covered by
test_gpu_properties.py and by bench.py's parity_sample): the engine runs the whole table, the oracle re-runs >= 500 sampled
targets on the engine's own inputs (neighbour lists; for Fisher-z the engine's Float32 cor_mat restricted to the variables
involved) and must reproduce PC sets, statistics (bit-exact for fz / fz_nz, 1e-12 for mi), p-values and test counts."""
import numpy as np
import pytest

import fwload
from oracle import fwo
from oracle import parity as opar

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fw():
    return fwload.load()


@pytest.fixture(scope="module")
def synth():
    return fwload.load_sub("synth")


def _sample(p, k, seed):
    return np.sort(np.random.default_rng(seed).choice(p, size=k, replace=False))


def test_c2_shape_fz(fw, synth):
    p, n = 10000, 2000
    x = synth.clique(p, n, B=24, seed=synth.BASE_SEED + 1)
    eng = fw.Engine(0)
    eng.set_data_colmajor(x, "fz")
    eng.pairwise_prefetch(0.01, 20)                          # pairwise candidates from the GEMM epilogue
    eng.cor(want_host=False)
    uni = eng.pw_univar_neighbors(alpha=0.01, n_obs_min=20)
    st1 = eng.pairwise_stats()
    # the same stage without the epilogue collection (one scan of the resident matrix): identical lists
    eng.pairwise_prefetch(0.0, 0)
    eng.cor(want_host=False)
    uni2 = eng.pw_univar_neighbors(alpha=0.01, n_obs_min=20)
    assert (uni.offsets == uni2.offsets).all() and (uni.nbr == uni2.nbr).all() and (uni.stat == uni2.stat).all() and (uni.pval == uni2.pval).all()
    assert st1 == eng.pairwise_stats()
    # whole pairwise stage against the oracle on the engine's cor_mat (5e7 pairs)
    ora = fwo.Oracle(np.zeros((n, p), np.float32), "fz")     # only the shape matters: the correlations come from set_cor
    ora.set_cor(eng.cor().astype(np.float64))
    off, nbr, st, ap = ora.pairwise(alpha=0.01, n_obs_min=20)
    assert (uni.offsets == off).all() and (uni.nbr == nbr).all() and (uni.stat == st).all()
    assert np.allclose(uni.pval, ap, rtol=1e-12, atol=1e-300)
    del ora
    order = fw.target_order(uni)
    res = eng.si_HITON_PC(order, max_k=3, alpha=0.01, n_obs_min=20, want_tpc=False)
    r = opar.sampled_hiton_parity(eng, "fz", x, _sample(p, 600, 1), res, uni, 3, 0.01, 20)
    assert r["targets"] == 600 and r["mismatches"] == 0, r
    assert r["max_stat_diff"] == 0.0 and r["cond_tests"] > 600 * 40000


def test_c3_shape_mi(fw, synth):
    p, n = 10000, 2000
    x = synth.binarize(synth.clique(p, n, B=24, seed=synth.BASE_SEED + 2))
    eng = fw.Engine(0)
    eng.set_data_colmajor(x, "mi")
    nom = fw.auto_n_obs_min("mi", 3, 5, max_level=2)
    uni = eng.pw_univar_neighbors(alpha=0.01, n_obs_min=nom)
    order = fw.target_order(uni)
    res = eng.si_HITON_PC(order, max_k=3, alpha=0.01, n_obs_min=nom, want_tpc=False)
    r = opar.sampled_hiton_parity(eng, "mi", x, _sample(p, 500, 2), res, uni, 3, 0.01, nom)
    assert r["targets"] == 500 and r["mismatches"] == 0, r
    assert r["max_stat_diff"] <= 1e-12 and r["cond_tests"] > 500 * 20000


def test_c5_shape_fznz(fw, synth):
    # C5-shaped: heterogeneous table with habitats, dropout and meta variables at the full 10 000 samples, reduced number of OTUs
    p, n = 4810, 10000
    x, meta_mask = synth.hetero(p, n, B=24, seed=synth.BASE_SEED + 4)
    eng = fw.Engine(0)
    eng.set_data_colmajor(x, "fz_nz")
    uni = eng.pw_univar_neighbors(alpha=0.01, n_obs_min=20)
    order = fw.target_order(uni)
    res = eng.si_HITON_PC(order, max_k=3, alpha=0.01, n_obs_min=20, want_tpc=False)
    pos = np.concatenate([_sample(p - 10, 500, 3), np.nonzero(np.isin(order, np.arange(p - 10, p)))[0]])    # + the 10 meta variables
    r = opar.sampled_hiton_parity(eng, "fz_nz", x, np.unique(pos), res, uni, 3, 0.01, 20, n_threads=16)
    assert r["targets"] >= 500 and r["mismatches"] == 0, r
    assert r["max_stat_diff"] == 0.0
