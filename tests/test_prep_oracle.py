"""The normalisation oracle (oracle/prep.py, numpy) against the reference's own fixtures: the six expected tables of
test/data/preprocessing_expected/ computed from test/data/HMP_SRA_gut/HMP_SRA_gut_small.tsv (test/preprocessing.jl:48-84),
plus the reference's behavioural tests (clr_adapt eps :37-45, zero-count filters :87-135)."""
import os

import numpy as np
import pytest

from oracle import prep

MODES = ["clr-adapt", "clr-nonzero", "clr-nonzero-binned", "pres-abs", "tss", "tss-nonzero-binned"]


@pytest.fixture(scope="module")
def fx(golden_dir):
    return np.load(os.path.join(golden_dir, "prep_fixtures.npz"))


@pytest.mark.parametrize("mode", MODES)
def test_expected_tables(fx, mode):
    data, cm, rm = prep.normalize(fx["counts"], norm_mode=mode)
    exp = fx[mode]
    if "binned" in mode:                      # legacy bin filtering of the fixtures (test/preprocessing.jl:70-74)
        data = data[:, [len(np.unique(data[:, j])) == 3 for j in range(data.shape[1])]]
    assert data.shape == exp.shape == (346, 50) and rm.sum() == 346 and cm.all()
    if exp.dtype.kind == "i":
        assert (data == exp).all()
    else:
        assert np.allclose(data, exp, rtol=1e-6, atol=1e-6)      # the fixtures are Float32-rounded text


def test_test_name_defaults(fx):
    for tn, mode in prep.DEFAULT_NORM.items():
        a = prep.normalize(fx["counts"], test_name=tn)[0]
        inv = {v: k for k, v in prep.MODE_MAP.items()}
        b = prep.normalize(fx["counts"], norm_mode=inv[mode])[0]
        assert a.dtype == b.dtype and (a == b).all()


def test_clr_adapt_eps():
    """test/preprocessing.jl:37-45: a sample whose adaptive pseudo-count underflows is removed, the rest stays finite."""
    s1 = np.concatenate([np.full(10000, 10000.0), np.zeros(10)])
    s2 = np.concatenate([np.full(10, 100.0), np.zeros(10000)])
    s3 = np.arange(1, 10011, dtype=np.float64)
    data, cm, rm = prep.normalize(np.stack([s1, s2, s3]), test_name="fz")
    assert np.isfinite(data).all() and data.shape[0] == 2


def test_zero_count_filters(fx):
    """test/preprocessing.jl:87-135: all-zero variables / samples are dropped; variables without exactly n_bins - 1 non-zero
    levels are dropped by the binned modes only."""
    data = fx["counts"].astype(np.float64)
    n, p = data.shape
    binfilt = np.vstack([np.zeros((n - 1, 10)), np.ones((1, 10))])
    rm = np.hstack([data, binfilt, np.zeros((n, 20))])
    rm = np.vstack([rm, np.zeros((10, rm.shape[1]))])
    for tn in ("mi", "mi_nz", "fz", "fz_nz"):
        out, cm, rmask = prep.normalize(rm, test_name=tn)
        zero_otus = 20 + (10 if tn == "mi_nz" else 0)
        assert out.shape[1] == rm.shape[1] - zero_otus, (tn, out.shape)
        assert out.shape[0] == rm.shape[0] - 15, (tn, out.shape)
        assert cm[: p].all() and not cm[p + 10:].any()
