"""Host-side meta-variable handling (flashweave.jl_b200/meta.py) against the reference's fixture
test/data/preprocessing_expected/meta_tiny_oneHotTest.tsv and the assertions of test/preprocessing.jl:144-185."""
import json
import os

import numpy as np
import pytest

import fwload

NORM = {"mi": "binary", "mi_nz": "binned_nz_clr", "fz": "clr_adapt", "fz_nz": "clr_nz"}      # preprocessing.jl:569-573


@pytest.fixture(scope="module")
def meta():
    return fwload.load_sub("meta")


@pytest.fixture(scope="module")
def fx(golden_dir):
    return json.load(open(os.path.join(golden_dir, "meta_onehot.json")))


def test_onehot_fixture(meta, fx):
    mat, names = meta.onehot(fx["columns"], fx["header"])
    exp = np.array(fx["expected"])
    assert names == fx["expected_header"]
    assert (mat[:, :-1] == exp[:, :-1]).all() and np.allclose(mat[:, -1], exp[:, -1], rtol=0, atol=0)
    # factors with two categories become integer codes, not indicator pairs; numeric columns pass through
    assert meta.check_onehot(["a", "b", "a"]) == (False, ["a", "b"])
    assert meta.factors_to_ints(["b", "a", "b"]) == [2, 1, 2]
    assert meta.check_onehot([1, 0, 2]) == (False, [])
    cols, nm = meta.onehot_column(["x", "y", "z", "x"], "F")
    assert nm == ["F_x", "F_y", "F_z"] and cols == [[1, 0, 0, 1], [0, 1, 0, 0], [0, 0, 1, 0]]


@pytest.mark.parametrize("test_name", ["fz", "mi", "fz_nz", "mi_nz"])
def test_meta_branch_of_preprocess_data(meta, fx, test_name):
    """test/preprocessing.jl:153-172: the encoded indicator columns come through every normalisation unchanged (shifted by +1 in
    the zero-ignoring Fisher-z mode), the continuous meta variable is binned to two levels for the discrete kinds."""
    exp = np.array(fx["expected"])
    keep_rows = np.ones(exp.shape[0], bool)
    keep_rows[[2, 11]] = False                          # a sample filter as the OTU table would produce
    mat, names = meta.prepare_meta(fx["columns"], fx["header"], NORM[test_name], obs_filter_mask=keep_rows)
    assert names == fx["expected_header"]
    A = mat[:, :-1].copy()
    if test_name == "fz_nz":
        A -= 1
    assert (A == exp[keep_rows][:, :-1]).all()
    if test_name.startswith("mi"):
        assert len(np.unique(mat[:, -1])) == 2
    else:
        assert np.allclose(mat[:, -1] - (1 if test_name == "fz_nz" and (exp[keep_rows][:, -1] == 0).any() else 0), exp[keep_rows][:, -1])
    # make_onehot = false: one column per meta variable
    mat2, names2 = meta.prepare_meta(fx["columns"], fx["header"], NORM[test_name], obs_filter_mask=keep_rows, make_onehot=False)
    assert mat2.shape[1] == len(fx["header"]) and names2 == fx["header"]


def test_iscontinuous_and_discretize(meta):
    assert not meta.iscontinuous([0, 1, 1, 0]) and not meta.iscontinuous([1.0, 0.0])
    assert meta.iscontinuous([0, 1, 2]) and meta.iscontinuous([0.5, 0.1]) and meta.iscontinuous([0, 3])
    assert list(meta.discretize([0.3, 0.1, 0.2, 0.4], 2)) == [1, 0, 0, 1]
    assert list(meta.discretize([5.0, 5.0, 1.0], 2)) == [1, 1, 0]              # tied ranks 2.5, 2.5, 1 -> / 2.5 -> 1, 1, 0.4
    # zero-variance meta variables are dropped
    mat, names = meta.prepare_meta([[1, 1, 1], [0, 1, 0]], ["c", "v"], "clr_adapt")
    assert names == ["v"] and mat.shape == (3, 1)
