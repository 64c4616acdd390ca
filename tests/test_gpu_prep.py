"""GPU parity tests of the normalisation step (fw_normalize_f32, csrc/prep.cuh) through the C ABI against the numpy oracle
(oracle/prep.py) and the reference's fixtures (tests/golden/prep_fixtures.npz, from test/data/preprocessing_expected/)."""
import json
import os

import numpy as np
import pytest

import fwload
from oracle import prep

pytestmark = pytest.mark.gpu
MODES = ["clr-adapt", "clr-nonzero", "clr-nonzero-binned", "pres-abs", "tss", "tss-nonzero-binned"]


@pytest.fixture(scope="module")
def fw():
    return fwload.load()


@pytest.fixture(scope="module")
def fx(golden_dir):
    return np.load(os.path.join(golden_dir, "prep_fixtures.npz"))


def _counts(n, p, seed, zero_frac=0.5):
    rng = np.random.default_rng(seed)
    depth = rng.lognormal(8.0, 1.0, size=(n, 1))
    comp = rng.dirichlet(np.full(p, 0.3), size=1)
    lam = depth * comp * (rng.random((n, p)) > zero_frac)
    x = rng.poisson(lam).astype(np.float64)
    x[:, 3] = 0                       # never observed
    x[:, 5] = 7                       # constant: zero variance
    x[11] = 0                         # sample without reads
    return x


def _check(mode, got, want):
    gd, wd = got["data"], want[0]
    assert (got["col_mask"] == want[1]).all() and (got["obs_filter_mask"] == want[2]).all(), mode
    assert gd.shape == wd.shape, (mode, gd.shape, wd.shape)
    if wd.dtype.kind == "i":
        assert gd.dtype == np.int32
        assert (gd != wd).sum() <= 2, (mode, int((gd != wd).sum()))          # a rank tie broken by the last ulp of log
    else:
        assert gd.dtype == np.float32
        assert np.allclose(gd, wd, rtol=2e-6, atol=1e-6), (mode, float(np.abs(gd - wd).max()))
        assert (gd == wd).mean() > 0.99, (mode, float((gd == wd).mean()))     # device vs host log: last-ulp flips after Float32 rounding only


@pytest.mark.parametrize("mode", MODES)
def test_reference_fixtures(fw, fx, mode):
    eng = fw.Engine(0)
    got = eng.normalize_data(fx["counts"], norm_mode=mode)
    data = got["data"]
    exp = fx[mode]
    if "binned" in mode:
        data = data[:, [len(np.unique(data[:, j])) == 3 for j in range(data.shape[1])]]
    assert data.shape == exp.shape == (346, 50) and got["obs_filter_mask"].sum() == 346 and got["col_mask"].all()
    if exp.dtype.kind == "i":
        assert (data == exp).all()
    else:
        assert np.allclose(data, exp, rtol=1e-6, atol=1e-6)
    _check(mode, got, prep.normalize(fx["counts"], norm_mode=mode))


@pytest.mark.parametrize("mode", MODES)
def test_random_counts(fw, mode):
    x = _counts(1500, 260, seed=3)
    eng = fw.Engine(0)
    _check(mode, eng.normalize_data(x, norm_mode=mode), prep.normalize(x, norm_mode=mode))


def test_behaviour_of_the_reference_tests(fw, fx):
    eng = fw.Engine(0)
    # test/preprocessing.jl:37-45 (clr_adapt eps)
    s1 = np.concatenate([np.full(10000, 10000.0), np.zeros(10)])
    s2 = np.concatenate([np.full(10, 100.0), np.zeros(10000)])
    s3 = np.arange(1, 10011, dtype=np.float64)
    got = eng.normalize_data(np.stack([s1, s2, s3]), test_name="fz")
    assert np.isfinite(got["data"]).all() and got["data"].shape[0] == 2
    # test/preprocessing.jl:87-135 (zero-count variables / samples)
    data = fx["counts"].astype(np.float64)
    n, p = data.shape
    rm = np.hstack([data, np.vstack([np.zeros((n - 1, 10)), np.ones((1, 10))]), np.zeros((n, 20))])
    rm = np.vstack([rm, np.zeros((10, rm.shape[1]))])
    for tn in ("mi", "mi_nz", "fz", "fz_nz"):
        got = eng.normalize_data(rm, test_name=tn)
        assert got["data"].shape == (rm.shape[0] - 15, rm.shape[1] - 20 - (10 if tn == "mi_nz" else 0)), (tn, got["data"].shape)
        _check(tn, got, prep.normalize(rm, test_name=tn))
    # nothing left / argument errors
    got = eng.normalize_data(np.zeros((5, 4)), test_name="fz")
    assert got["data"] is None and not got["col_mask"].any()
    with pytest.raises(ValueError):
        eng.normalize_data(data, test_name="fz", norm_mode="tss")


def test_counts_to_network(fw, fx, golden_dir):
    """raw counts -> normalisation on the device -> resident table -> pairwise stage / HITON-PC: the reference's expected graphs
    (test/data/learning_expected, computed by the reference from the same counts) come out of the whole chain."""
    graphs = json.load(open(os.path.join(golden_dir, "learning_expected.json")))
    for tn in ("fz", "mi", "fz_nz", "mi_nz"):
        eng = fw.Engine(0)
        got = eng.normalize_data(fx["counts"], test_name=tn, want_host=False)
        assert got["kind"] == tn and eng.n == 346
        r = eng.LGL(max_k=0)
        want = {(a, b) for a, b, _ in graphs[f"exp_{tn}_maxk0"]}
        if tn == "mi_nz":
            # the fixture tables keep the 2-level columns the current code filters out (test/preprocessing.jl:70-74): map indices
            keep = np.flatnonzero(got["col_mask"])
            have = {(int(keep[a]), int(keep[b])) for a, b, _ in r["edges"]}
        else:
            have = {(a, b) for a, b, _ in r["edges"]}
        assert have == want, (tn, len(have), len(want))


def test_sparse_csc_input(fw, fx):
    """SparseMatrixCSC triples (the reference's default make_sparse = true tables) give the same resident table, and therefore
    the same tests, as the dense upload; index base 1 as the Julia glue passes them; bad structure fails loudly."""
    import scipy.sparse as sp
    for kind, dense in (("fz_nz", prep.normalize(fx["counts"], test_name="fz_nz")[0]), ("mi_nz", prep.normalize(fx["counts"], test_name="mi_nz")[0])):
        n, p = dense.shape
        m = sp.csc_matrix(dense)
        a = fw.Engine(0).set_data(dense, kind)
        b = fw.Engine(0).set_data_csc(m.indptr, m.indices, m.data, n, p, kind)
        assert (b.get_data() == dense).all()
        ra = a.pw_univar_neighbors(alpha=0.01, n_obs_min=20)
        rb = b.pw_univar_neighbors(alpha=0.01, n_obs_min=20)
        assert (ra.offsets == rb.offsets).all() and (ra.nbr == rb.nbr).all() and (ra.stat == rb.stat).all()
        c = fw.Engine(0)
        c.L.fw_set_index_base(c.h, 1)
        c.set_data_csc(m.indptr + 1, m.indices + 1, m.data, n, p, kind)
        assert (c.get_data() == dense).all()
        c.L.fw_set_index_base(c.h, 0)
        with pytest.raises(fw.FwError):
            fw.Engine(0).set_data_csc(m.indptr, m.indices + n, m.data, n, p, kind)       # rows out of range
        with pytest.raises(fw.FwError):
            fw.Engine(0).set_data_csc(m.indptr + 1, m.indices, m.data, n, p, kind)       # colptr[0] != index base


def test_meta_variables(fw, golden_dir):
    """test/preprocessing.jl:144-185 through Engine.normalize_data: OTU counts normalised on the device, meta variables one-hot
    encoded / discretised / shifted on the host and appended; the combined table is the resident one."""
    fxm = json.load(open(os.path.join(golden_dir, "meta_onehot.json")))
    counts = np.array(fxm["counts"], dtype=np.float64)
    exp = np.array(fxm["expected"])
    meta = fwload.load_sub("meta")
    for tn in ("fz", "mi", "fz_nz", "mi_nz"):
        eng = fw.Engine(0)
        got = eng.normalize_data(counts, test_name=tn, meta_data=fxm["columns"], meta_header=fxm["header"])
        data, mm, rm = got["data"], got["meta_mask"], got["obs_filter_mask"]
        assert got["meta_names"] == fxm["expected_header"] and mm.sum() == len(fxm["expected_header"])
        A = data[:, mm][:, :-1].astype(np.float64)
        if tn == "fz_nz":
            A -= 1                                      # the +1 shift of one-hot variables in the zero-ignoring mode
        assert (A == exp[rm][:, :-1]).all(), tn
        if tn.startswith("mi"):
            assert len(np.unique(data[:, -1])) == 2
        # the same combination from the oracle's normalisation
        w = prep.normalize(counts, test_name=tn)
        inv = {v: k for k, v in prep.MODE_MAP.items()}
        wc, wm, wn = meta.combine_with_meta(w[0], w[2], fxm["columns"], fxm["header"], inv[prep.DEFAULT_NORM[tn]])
        assert data.shape == wc.shape and (mm == wm).all()
        assert np.allclose(data, wc, rtol=2e-6, atol=1e-6)
        # resident table = returned table; the hot path runs on it
        assert eng.p == data.shape[1] and eng.n == data.shape[0]
        assert (eng.get_data() == data).all()
        assert (eng.meta_mask() == mm).all()                  # fw_set_meta_mask: carried with the resident table ...
        r = eng.LGL(max_k=0)
        out = os.path.join(os.environ.get("TMPDIR", "/tmp"), "fw_meta_%s.edgelist" % tn)
        fw.write_edgelist(out, r["edges"], header=["v%d" % i for i in range(eng.p)], meta_mask=eng.meta_mask())
        assert open(out).read().split("\n")[1] == "# meta mask\t" + ",".join("true" if m else "false" for m in mm)   # ... to the `# meta mask` line (io.jl:338-346)
        eng.set_data(data[:, :5], tn)
        assert not eng.meta_mask().any()                      # a new table clears it
