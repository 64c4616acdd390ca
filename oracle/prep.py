"""CPU restatement (numpy) of the reference's normalisation step for dense OTU tables WITHOUT meta variables — the step in
front of the CI-test hot path.  TEST INFRASTRUCTURE ONLY: imported by tests/ (and nothing else); the
product path is the CUDA implementation behind fw_normalize_f32 (include/fwgpu.h).

Follows /root/reference/src/preprocessing.jl (paths below relative to that checkout):
  normalize_data                 src/preprocessing.jl:660-684   (mode names -> internal norm strings)
  preprocess_data                src/preprocessing.jl:412-563   (filter -> normalise -> level filter -> target precision)
  filter_by_variance             src/preprocessing.jl:367-409
  rownorm!                       src/preprocessing.jl:348
  clr! / adaptive_clr! / adaptive_pseudocount!   src/preprocessing.jl:133-215
  discretize / discretize_nz     src/preprocessing.jl:217-292   (tied ranks, "median" method)
  presabs_norm!                  src/preprocessing.jl:364-365
Pinned by the reference's own fixtures test/data/preprocessing_expected/*.tsv computed from
test/data/HMP_SRA_gut/HMP_SRA_gut_small.tsv (test/preprocessing.jl:48-84), see tests/test_prep_oracle.py.
"""
import numpy as np

MODE_MAP = {"clr-adapt": "clr_adapt", "clr-nonzero": "clr_nz", "clr-nonzero-binned": "binned_nz_clr", "pres-abs": "binary",
            "tss": "rows", "tss-nonzero-binned": "binned_nz_rows"}                       # preprocessing.jl:666-668
DEFAULT_NORM = {"mi": "binary", "mi_nz": "binned_nz_clr", "fz": "clr_adapt", "fz_nz": "clr_nz"}   # preprocessing.jl:569-573


def _tiedrank(v):
    """StatsBase.tiedrank: average of the 1-based positions of equal values."""
    order = np.argsort(v, kind="stable")
    sv = v[order]
    n = len(v)
    ranks = np.empty(n, np.float64)
    i = 0
    while i < n:
        j = i
        while j + 1 < n and sv[j + 1] == sv[i]:
            j += 1
        ranks[order[i:j + 1]] = 0.5 * ((i + 1) + (j + 1))
        i = j + 1
    return ranks


def _discretize(v, n_bins):
    """preprocessing.jl:238-253 (rank_method "tied", disc_method "median")."""
    if len(v) == 0:
        return np.zeros(0, np.int64)
    r = _tiedrank(v)
    r = r / r.max()
    step = (1.0 / n_bins) + 1e-5
    return np.floor(r / step).astype(np.int64)


def _clr_nz(x64):
    """clr!(X; pseudo_count=0.0, ignore_zeros=true), preprocessing.jl:192-207."""
    out = np.zeros_like(x64)
    for i in range(x64.shape[0]):
        nzm = x64[i] != 0
        if nzm.any():
            g = np.exp(np.mean(np.log(x64[i][nzm])))
            out[i][nzm] = np.log(x64[i][nzm] / g)
    return out


def normalize(counts, norm_mode="", test_name="", n_bins=3):
    """counts: [n samples, p variables].  Returns (data [n', p'], col_mask[p], row_mask[n]) with data float32 for the continuous
    norms and int32 for the discrete ones (convert_to_target_prec, prec = 32)."""
    assert bool(norm_mode) != bool(test_name)
    norm = MODE_MAP[norm_mode] if norm_mode else DEFAULT_NORM[test_name]
    x = np.asarray(counts).astype(np.float32)                               # check_convert_sparse: Matrix{Float32}
    n, p = x.shape
    col_mask = x.var(axis=0, ddof=1) > 0 if n > 1 else np.zeros(p, bool)    # filter_by_variance
    x = x[:, col_mask]
    row_mask = x.sum(axis=1) > 0
    x = x[row_mask]
    if norm == "rows":
        data = (x / x.sum(axis=1, keepdims=True, dtype=np.float32)).astype(np.float32)
    elif norm == "clr_nz":
        data = _clr_nz(x.astype(np.float64)).astype(np.float32)
    elif norm == "clr_adapt":
        x64 = x.astype(np.float64)
        depth = x64.sum(axis=1)
        md = x64[int(np.argmax(depth))]                                      # first maximum (findmax)
        min_ab = x64[x64 != 0].min()
        base = 1.0 if min_ab >= 1 else min_ab / 10
        pv = x64.shape[1]
        k = int((md == 0).sum()); nprod1 = float(np.log(md[md != 0]).sum())
        pcs = np.empty(x64.shape[0])
        for i in range(x64.shape[0]):
            s = x64[i]
            nz0 = int((s == 0).sum()); nprod2 = float(np.log(s[s != 0]).sum())
            pcs[i] = np.exp((1.0 / (nz0 - pv)) * ((k - pv) * np.log(base) + nprod1 - nprod2)) if nz0 != pv else np.nan
        keep = pcs != 0
        x64 = x64[keep]; pcs = pcs[keep]
        rm = np.flatnonzero(row_mask)[~keep]
        row_mask = row_mask.copy(); row_mask[rm] = False
        for i in range(x64.shape[0]):
            x64[i][x64[i] == 0] = pcs[i]
        g = np.exp(np.mean(np.log(x64), axis=1, keepdims=True))
        data = np.log(x64 / g).astype(np.float32)
    elif norm == "binary":
        data = np.sign(x).astype(np.int32)
        lv = np.array([len(np.unique(data[:, j])) for j in range(data.shape[1])]) == 2
        data = data[:, lv]
        cm = col_mask.copy(); cm[np.flatnonzero(col_mask)[~lv]] = False; col_mask = cm
    elif norm in ("binned_nz_clr", "binned_nz_rows"):
        nz_mask = x != 0
        if norm == "binned_nz_clr":
            v = _clr_nz(x.astype(np.float64))
        else:
            v = (x / x.sum(axis=1, keepdims=True, dtype=np.float32)).astype(np.float32).astype(np.float64)
        data = np.zeros(x.shape, np.int32)
        for j in range(x.shape[1]):
            m = nz_mask[:, j]
            if m.any():
                data[m, j] = _discretize(v[m, j], n_bins - 1) + 1
        lv = np.array([len(np.unique(data[:, j][data[:, j] != 0])) for j in range(data.shape[1])]) == n_bins - 1
        data = data[:, lv]
        cm = col_mask.copy(); cm[np.flatnonzero(col_mask)[~lv]] = False; col_mask = cm
    else:
        raise ValueError(norm)
    return data, col_mask, row_mask
