"""The pairwise stage of the table-based kinds split over ranks (fw_pairwise_partial / fw_pairwise_merge, include/fwgpu.h):
the ranks are emulated one after the other on ONE device - each evaluates the pairs of its X groups, the records are concatenated
(what parallel.allgather_records does between processes) and merged.  Neighbour lists, adjusted p-values and the counters must be
identical to fw_pairwise on one GPU (tests.jl:436-532, statfuns.jl:326-350), for every world size and both fz_nz back-ends."""
import os

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


def _check(fw, eng, kind, nom, worlds, **kw):
    want = eng.pw_univar_neighbors(alpha=0.01, n_obs_min=nom, kind=kind, **kw)
    st_want = eng.pairwise_stats()
    for world in worlds:
        recs = [eng.pairwise_partial(r, world, alpha=0.01, n_obs_min=nom, kind=kind, correct_reliable_only=kw.get("correct_reliable_only", True))
                for r in range(world)]
        # every pair belongs to exactly one rank
        keys = np.concatenate([r["x"].astype(np.int64) * eng.p + r["y"] for r in recs])
        assert len(np.unique(keys)) == len(keys) == st_want["n_raw_sig"]
        assert sum(r["n_reliable"] for r in recs) == st_want["n_reliable"] or not kw.get("correct_reliable_only", True)
        if world > 1 and eng.p > 2048:
            assert sum(len(r["x"]) > 0 for r in recs) > 1                     # the work really is split
        got = eng.pairwise_merge(recs[::-1], alpha=0.01, kind=kind, **kw)     # any order of the ranks
        assert (got.offsets == want.offsets).all() and (got.nbr == want.nbr).all()
        assert (got.stat == want.stat).all() and (got.pval == want.pval).all()
        st = eng.pairwise_stats()
        assert st["n_raw_sig"] == st_want["n_raw_sig"] and st["n_tests"] == st_want["n_tests"]


def test_sharded_pairwise_fznz(fw, synth):
    x = synth.hetero(3300, 600, B=12, seed=11)[0]                             # 4 X groups of 1024
    for tc in ("1", "0"):
        os.environ["FWGPU_FZNZ_TC"] = tc
        try:
            eng = fw.Engine(0)
            eng.set_data_colmajor(x, "fz_nz")
            _check(fw, eng, "fz_nz", 20, (1, 2, 3, 8))
            _check(fw, eng, "fz_nz", 20, (2,), correct_reliable_only=False)
        finally:
            os.environ.pop("FWGPU_FZNZ_TC", None)


def test_sharded_pairwise_discrete(fw, synth, golden_dir):
    xb = synth.binarize(synth.clique(2500, 400, B=10, seed=12))
    eng = fw.Engine(0)
    eng.set_data_colmajor(xb, "mi")
    _check(fw, eng, "mi", fw.auto_n_obs_min("mi", 3, 5, max_level=2), (1, 2, 4))
    hmp = np.load(os.path.join(golden_dir, "hmp_inputs.npz"))
    A = np.ascontiguousarray(np.array(hmp["mi_nz"], dtype=np.int32).T)
    eng = fw.Engine(0)
    eng.set_data_colmajor(A, "mi_nz")
    _check(fw, eng, "mi_nz", fw.auto_n_obs_min("mi_nz", 3, 5, max_level=3), (1, 2))


def test_sharded_pairwise_errors(fw, synth):
    eng = fw.Engine(0)
    x = synth.clique(64, 200, B=8, seed=1)
    eng.set_data_colmajor(x, "fz")
    eng.cor(want_host=False)
    with pytest.raises(fw.FwError):
        eng.pairwise_partial(0, 2, kind="fz")                                # fz shares its pairwise stage through the group path
    eng.set_data_colmajor(x, "fz_nz")
    with pytest.raises(fw.FwError):
        eng.pairwise_partial(2, 2, kind="fz_nz")
