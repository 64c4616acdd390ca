"""Host-side handling of meta variables in front of the normalisation step (pure host logic, no compute kernels):
string factors -> integers, one-hot encoding of factors with more than two categories, discretisation of continuous meta
variables for the discrete test kinds, the "+1 shift" of the zero-ignoring Fisher-z mode, and the zero-variance filter.

Mirrors /root/reference/src/preprocessing.jl (paths relative to that checkout):
  factors_to_ints                 src/preprocessing.jl:42-56
  check_onehot / onehot           src/preprocessing.jl:59-118
  iscontinuous / discretize_meta! src/preprocessing.jl:294-315
  discretize ("median", tied ranks)  src/preprocessing.jl:238-253
  the meta branch of preprocess_data src/preprocessing.jl:418-446, 523-556
Pinned by the reference's fixture test/data/preprocessing_expected/meta_tiny_oneHotTest.tsv (tests/test_meta_host.py).
"""
import numpy as np


def _is_string_column(x):
    return isinstance(x[0], str)


def factors_to_ints(x):
    """String factors -> 1-based integer codes in sorted category order; numeric vectors pass through."""
    if _is_string_column(x):
        cats = sorted(set(x))
        fmap = {c: i + 1 for i, c in enumerate(cats)}
        return [fmap[v] for v in x]
    return list(x)


def check_onehot(x):
    """(needs one-hot?, categories): numeric vectors never; string factors iff more than two categories."""
    if not _is_string_column(x):
        return False, []
    cats = sorted(set(x))
    return len(cats) > 2, cats


def onehot_column(x, var_name="", check=True):
    needs, cats = check_onehot(x)
    if not check or needs:
        if not cats:
            cats = sorted(set(x))
        cols = [[1 if v == c else 0 for v in x] for c in cats]
        names = [var_name + "_" + str(c) for c in cats] if var_name else []
    else:
        cols, names = [factors_to_ints(x)], [var_name]
    return cols, names


def onehot(columns, names=None, check=True):
    """columns: list of meta variables (each a list of numbers or strings).  Returns ([n, q] float64 matrix, names)."""
    names = list(names) if names else []
    out_cols, out_names = [], []
    for i, col in enumerate(columns):
        c, nm = onehot_column(list(col), names[i] if names else "", check)
        out_cols += c
        out_names += nm
    mat = np.array(out_cols, dtype=np.float64).T if out_cols else np.zeros((0, 0))
    return mat, (out_names if names else [])


def _tiedrank(v):
    order = np.argsort(v, kind="stable")
    sv = np.asarray(v)[order]
    r = np.empty(len(v), np.float64)
    i = 0
    while i < len(v):
        j = i
        while j + 1 < len(v) and sv[j + 1] == sv[i]:
            j += 1
        r[order[i:j + 1]] = 0.5 * ((i + 1) + (j + 1))
        i = j + 1
    return r


def discretize(v, n_bins=3):
    """rank_method "tied", disc_method "median" (preprocessing.jl:238-253)."""
    v = np.asarray(v, dtype=np.float64)
    if len(v) == 0:
        return v.astype(np.int64)
    r = _tiedrank(v)
    r = r / r.max()
    return np.floor(r / ((1.0 / n_bins) + 1e-5)).astype(np.int64)


def iscontinuous(v):
    """preprocessing.jl:296-303: integer-valued vectors count as continuous unless they look like a binary indicator."""
    v = np.asarray(v, dtype=np.float64)
    # isapprox(round.(x), x): norm(x - y) <= sqrt(eps) * max(norm(x), norm(y))   (Julia's array isapprox)
    rv = np.round(v)
    if np.linalg.norm(rv - v) <= np.sqrt(np.finfo(np.float64).eps) * max(np.linalg.norm(rv), np.linalg.norm(v)):
        return bool(v.max() > 1 or len(np.unique(v)) > 2)
    return True


def prepare_meta(columns, names, norm, obs_filter_mask=None, make_onehot=True, n_bins=2):
    """The meta branch of preprocess_data for one of the internal norm names ("clr_adapt", "clr_nz", "binary", "binned_nz_clr",
    "rows", "binned_nz_rows").  Returns ([n', q'] float64, names): one-hot / integer encoding, the sample filter of the OTU table,
    discretisation of continuous meta variables for the non-continuous norms, +1 shift of variables that contain zeros for
    "clr_nz" (zeros mean "absent" to the zero-ignoring test), removal of zero-variance variables."""
    if make_onehot:
        mat, names = onehot(columns, names)
    else:
        mat = np.array([factors_to_ints(list(c)) for c in columns], dtype=np.float64).T
        names = list(names) if names else []
    if obs_filter_mask is not None:
        mat = mat[np.asarray(obs_filter_mask, bool)]
    continuous_norm = norm == "rows" or norm.startswith("clr")
    if not continuous_norm:
        for j in range(mat.shape[1]):
            if iscontinuous(mat[:, j]):
                mat[:, j] = discretize(mat[:, j], n_bins)
    if norm == "clr_nz":
        for j in range(mat.shape[1]):
            if (mat[:, j] == 0).any():
                mat[:, j] += 1
    keep = mat.var(axis=0, ddof=1) > 0 if mat.shape[0] > 1 else np.zeros(mat.shape[1], bool)
    mat = mat[:, keep]
    if names:
        names = [nm for nm, k in zip(names, keep) if k]
    return mat, names


INTERNAL_NORM = {"tss": "rows", "clr-adapt": "clr_adapt", "clr-nonzero": "clr_nz", "pres-abs": "binary",
                 "clr-nonzero-binned": "binned_nz_clr", "tss-nonzero-binned": "binned_nz_rows"}          # preprocessing.jl:666-668


def combine_with_meta(norm_table, obs_filter_mask, meta_columns, meta_header, norm_mode, make_onehot=True):
    """hcat(data, meta_data) of preprocess_data (preprocessing.jl:549-556): the normalised OTU table [n', p'] followed by the
    prepared meta variables, in the table's element type (Float32 / Int32).  Returns (combined [n', p' + q'], meta_mask, meta_names)."""
    mat, names = prepare_meta(meta_columns, meta_header, INTERNAL_NORM[norm_mode], obs_filter_mask=obs_filter_mask, make_onehot=make_onehot)
    norm_table = np.asarray(norm_table)
    if mat.shape[0] != norm_table.shape[0]:
        raise ValueError("meta data has %d samples after filtering, the table %d" % (mat.shape[0], norm_table.shape[0]))
    combined = np.concatenate([norm_table, mat.astype(norm_table.dtype)], axis=1)
    meta_mask = np.concatenate([np.zeros(norm_table.shape[1], bool), np.ones(mat.shape[1], bool)])
    return combined, meta_mask, names
