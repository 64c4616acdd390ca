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
