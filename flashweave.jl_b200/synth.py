"""Synthetic OTU tables of SURVEY.md §8d, shared by the oracle and the GPU path.

All generators return a C-contiguous [p, n] float32 array (= the column-major n x p
Matrix{Float32} the reference works on, one variable per row here) in already
"normalised" form (normalize=false semantics).  Seed = 20190802 + config_index.
"""
import numpy as np

BASE_SEED = 20190802


def clique(p, n, B=24, seed=BASE_SEED, dtype=np.float32):
    """latent "clique-B": blocks of B variables sharing one factor, x_i = 0.8 f_b + 0.6 e_i
    (pairwise r = 0.64, partial r given 3 block-mates ~ 0.22): nothing exits early."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = np.empty((p, n), dtype)
    nb = (p + B - 1) // B
    for b in range(nb):
        lo, hi = b * B, min((b + 1) * B, p)
        f = rng.standard_normal(n, dtype=np.float32)
        e = rng.standard_normal((hi - lo, n), dtype=np.float32)
        out[lo:hi] = 0.8 * f[None, :] + 0.6 * e
    return out


def chain(p, n, B=32, rho=0.7, seed=BASE_SEED, dtype=np.float32):
    """latent "chain": blocks of B variables, AR(1) inside a block; true skeleton = chain edges."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = np.empty((p, n), dtype)
    s = np.float32(np.sqrt(1.0 - rho * rho))
    for v in range(p):
        e = rng.standard_normal(n, dtype=np.float32)
        if v % B == 0:
            out[v] = e
        else:
            out[v] = np.float32(rho) * out[v - 1] + s * e
    return out


def binarize(x_pn):
    """C3: uint8(latent > column median) -> presence/absence codes (levels 2, max_val 1)."""
    med = np.median(x_pn, axis=1, keepdims=True)
    return (x_pn > med).astype(np.int32)


def three_level(x_pn, zero_frac=0.4, seed=0):
    """mi_nz-style codes: a fraction of structural zeros (absent), non-zeros binned to {1, 2} at the
    per-variable median of the non-zero entries (the reference's clr_nonzero_binned form)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    p, n = x_pn.shape
    out = np.zeros((p, n), np.int32)
    present = rng.random((p, n)) >= zero_frac
    for v in range(p):
        nzv = x_pn[v][present[v]]
        if nzv.size == 0:
            continue
        med = np.median(nzv)
        out[v][present[v]] = 1 + (nzv > med)
    return out


def with_zeros(x_pn, zero_frac=0.4, seed=0):
    """fz_nz-style table: structural zeros (absent taxa) punched into a continuous table; non-zeros keep their value."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = x_pn.copy()
    out[rng.random(x_pn.shape) < zero_frac] = 0.0
    return out


def hetero(p, n, B=24, H=8, n_meta=10, dropout=0.1, seed=BASE_SEED + 4):
    """C5: heterogeneous fz_nz table + meta variables.  H habitats, sample -> habitat uniform; every
    block of B OTUs is present in a habitat w.p. 0.5 (at least one); absent => structural 0.0, plus `dropout` random
    zeros; non-zero entries keep the clique-B latent value.  The last n_meta rows are meta variables that are never
    zero: H habitat indicators coded {1, 2} (the reference's "+1 shift", preprocessing.jl:537-545) and the rest
    continuous N(0,1) + 0.5 * (block-0 latent factor).  Returns ([p, n] float32, meta_mask[p] bool)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n_otu = p - n_meta
    out = np.empty((p, n), np.float32)
    hab = rng.integers(0, H, size=n)
    nb = (n_otu + B - 1) // B
    f0 = None
    for b in range(nb):
        lo, hi = b * B, min((b + 1) * B, n_otu)
        f = rng.standard_normal(n, dtype=np.float32)
        if b == 0:
            f0 = f
        e = rng.standard_normal((hi - lo, n), dtype=np.float32)
        blk = 0.8 * f[None, :] + 0.6 * e
        blk[blk == 0.0] = 1e-3
        pres = rng.random(H) < 0.5
        if not pres.any():
            pres[rng.integers(0, H)] = True
        blk[:, ~pres[hab]] = 0.0
        blk[rng.random(blk.shape) < dropout] = 0.0
        out[lo:hi] = blk
    for h in range(min(H, n_meta)):
        out[n_otu + h] = 1.0 + (hab == h)
    for j in range(min(H, n_meta), n_meta):
        out[n_otu + j] = rng.standard_normal(n, dtype=np.float32) + 0.5 * f0
    mask = np.zeros(p, bool)
    mask[n_otu:] = True
    return out, mask
