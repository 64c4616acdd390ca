"""ctypes binding of liboracle.so — the CPU ORACLE (test infrastructure only).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  The product path (flashweave.jl_b200, libfwgpu.so) never does.

All indices are 0-based (the reference is 1-based Julia).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

KINDS = {"mi": 0, "mi_nz": 1, "fz": 2, "fz_nz": 3}


class Result(C.Structure):
    # src/types.jl:140-145
    _fields_ = [("stat", C.c_double), ("pval", C.c_double), ("df", C.c_int64),
                ("suff_power", C.c_uint8), ("pad", C.c_uint8 * 7)]

    def astuple(self):
        return (self.stat, self.pval, int(self.df), bool(self.suff_power))


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "fw_oracle.cpp")
    if force or not os.path.exists(so) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(so)):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = build()
        L = C.CDLL(so)
        i64, dbl, p = C.c_int64, C.c_double, C.c_void_p
        L.fwo_create.restype = p
        L.fwo_create.argtypes = [C.c_int, i64, i64, p, p, C.c_int]
        L.fwo_destroy.argtypes = [p]
        L.fwo_get_levels.argtypes = [p, p, p]
        L.fwo_set_sparse_semantics.argtypes = [p, C.c_int]
        L.fwo_compute_cor.argtypes = [p]
        L.fwo_set_cor.argtypes = [p, p]
        L.fwo_get_cor.argtypes = [p, p]
        L.fwo_alloc_scratch_cor.argtypes = [p]
        L.fwo_fz_pval.restype = dbl
        L.fwo_fz_pval.argtypes = [dbl, i64, i64]
        L.fwo_chisq_sf.restype = dbl
        L.fwo_chisq_sf.argtypes = [i64, dbl]
        L.fwo_mi_pval.restype = dbl
        L.fwo_mi_pval.argtypes = [dbl, i64, i64]
        L.fwo_benjamini_hochberg.argtypes = [p, i64, dbl, i64]
        L.fwo_pcor_rec.restype = dbl
        L.fwo_pcor_rec.argtypes = [p, i64, C.c_int, i64, i64, p, C.c_int, p]
        L.fwo_mutual_information.restype = dbl
        L.fwo_mutual_information.argtypes = [p, i64, i64, i64]
        L.fwo_test_uni.argtypes = [p, i64, p, i64, i64, i64, C.c_int, p, i64, p]
        L.fwo_test_cond.argtypes = [p, i64, i64, p, C.c_int, i64, i64, C.c_int, C.c_int, p, i64, p, p, p]
        L.fwo_test_subsets.argtypes = [p, i64, i64, p, i64, C.c_int, dbl, i64, i64, i64, C.c_int, p, i64, p, p, p, p, p]
        L.fwo_pairwise.restype = i64
        L.fwo_pairwise.argtypes = [p, dbl, i64, i64, C.c_int, C.c_int, p, p, p, p, p, p]
        L.fwo_auto_n_obs_min.restype = i64
        L.fwo_auto_n_obs_min.argtypes = [p, C.c_int, i64]
        L.fwo_hiton_pc.restype = i64
        L.fwo_hiton_pc.argtypes = [p, i64, p, p, p, i64, C.c_int, dbl, i64, i64, i64, p, i64, p, p, p, p]
        L.fwo_hiton_pc_batch.restype = None
        L.fwo_hiton_pc_batch.argtypes = [p, i64, p, p, p, p, p, C.c_int, dbl, i64, i64, i64, C.c_int, p, p, p, p, p]
        L.fwo_lgl.restype = i64
        L.fwo_lgl.argtypes = [p, C.c_int, dbl, i64, i64, i64, C.c_int, C.c_int, C.c_int, p, i64,
                              p, p, p, i64, p, p, p, p, p, p, i64, p]
        L.fwo_num_threads.restype = C.c_int
        _LIB = L
    return _LIB


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _i64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int64))


class Oracle:
    """One reference `test_obj` + data (make_test_object, src/misc.jl:34-45)."""

    def __init__(self, data, kind, cont32=True):
        self.L = lib()
        self.kind = kind
        k = KINDS[kind]
        data = np.asarray(data)
        self.n, self.p = data.shape
        if k < 2:
            self._d = np.asfortranarray(data.astype(np.int32))
            self.h = self.L.fwo_create(k, self.n, self.p, None, _ptr(self._d), int(cont32))
        else:
            self._d = np.asfortranarray(data.astype(np.float64))
            self.h = self.L.fwo_create(k, self.n, self.p, _ptr(self._d), None, int(cont32))
        self.cont32 = cont32

    def __del__(self):
        try:
            self.L.fwo_destroy(self.h)
        except Exception:
            pass

    def set_sparse_semantics(self, on=True):
        """discrete kinds: the reference's sparse-input code path (contingency.jl:80-480), its default for sensitive=false"""
        self.L.fwo_set_sparse_semantics(self.h, int(on))

    # -- precompute ---------------------------------------------------------------
    def levels(self):
        lv = np.zeros(self.p, np.int32)
        mv = np.zeros(self.p, np.int32)
        self.L.fwo_get_levels(self.h, _ptr(lv), _ptr(mv))
        return lv, mv

    def compute_cor(self):
        self.L.fwo_compute_cor(self.h)
        out = np.zeros((self.p, self.p), np.float64, order="F")
        self.L.fwo_get_cor(self.h, _ptr(out))
        return out

    def set_cor(self, cor):
        c = np.asfortranarray(np.asarray(cor, dtype=np.float64))
        assert c.shape == (self.p, self.p)
        self.L.fwo_set_cor(self.h, _ptr(c))

    def auto_n_obs_min(self, max_k, hps=5):
        return int(self.L.fwo_auto_n_obs_min(self.h, max_k, hps))

    # -- single tests -------------------------------------------------------------
    def test_uni(self, X, Ys, hps=5, n_obs_min=0, trim_x=True, rows=None):
        Ys = _i64(Ys)
        out = (Result * len(Ys))()
        r = None if rows is None else np.ascontiguousarray(rows, dtype=np.int32)
        self.L.fwo_test_uni(self.h, X, _ptr(Ys), len(Ys), hps, n_obs_min, int(trim_x), _ptr(r), -1 if r is None else len(r), out)
        return [o.astuple() for o in out]

    def test_cond(self, X, Y, Zs, hps=5, n_obs_min=0, max_k=3, trim_xy=True, rows=None, want_ctab=False):
        Zs = _i64(Zs)
        out = Result()
        lz = C.c_int64(0)
        r = None if rows is None else np.ascontiguousarray(rows, dtype=np.int32)
        ctab = None
        if want_ctab:
            _, mv = self.levels()
            Lv = int(mv.max()) + 1
            ctab = np.zeros((Lv, Lv, Lv ** max(max_k, len(Zs))), np.int64, order="F")
        self.L.fwo_test_cond(self.h, X, Y, _ptr(Zs), len(Zs), hps, n_obs_min, max_k, int(trim_xy), _ptr(r),
                             -1 if r is None else len(r), C.byref(out), C.byref(lz), _ptr(ctab))
        if want_ctab:
            return out.astuple(), int(lz.value), ctab
        return out.astuple()

    def test_subsets(self, X, Y, Z_total, max_k=3, alpha=0.01, hps=5, n_obs_min=0, max_tests=0, trim_xy=True, rows=None):
        Z = _i64(Z_total)
        out = Result()
        Zs = np.zeros(3, np.int64)
        k = C.c_int(0)
        nt = C.c_int64(0)
        fr = C.c_double(0)
        r = None if rows is None else np.ascontiguousarray(rows, dtype=np.int32)
        self.L.fwo_test_subsets(self.h, X, Y, _ptr(Z), len(Z), max_k, alpha, hps, n_obs_min, max_tests, int(trim_xy),
                                _ptr(r), -1 if r is None else len(r), C.byref(out), _ptr(Zs), C.byref(k), C.byref(nt), C.byref(fr))
        return out.astuple(), tuple(int(z) for z in Zs[:k.value]), int(nt.value), fr.value

    # -- batches -------------------------------------------------------------------
    def pairwise(self, alpha=0.01, hps=5, n_obs_min=0, fdr=True, correct_reliable_only=True, want_raw=False):
        p = self.p
        off = np.zeros(p + 1, np.int64)
        npairs = p * (p - 1) // 2
        rs = np.zeros(npairs) if want_raw else None
        rp = np.zeros(npairs) if want_raw else None
        tot = self.L.fwo_pairwise(self.h, alpha, hps, n_obs_min, int(fdr), int(correct_reliable_only), _ptr(off), None, None, None, None, None)
        nbr = np.zeros(tot, np.int64)
        st = np.zeros(tot)
        ap = np.zeros(tot)
        self.L.fwo_pairwise(self.h, alpha, hps, n_obs_min, int(fdr), int(correct_reliable_only), _ptr(off), _ptr(nbr), _ptr(st), _ptr(ap), _ptr(rs), _ptr(rp))
        if want_raw:
            return off, nbr, st, ap, rs, rp
        return off, nbr, st, ap

    def hiton_pc(self, T, uni_nbr, uni_stat, uni_p, max_k=3, alpha=0.01, hps=5, n_obs_min=0, max_tests=0, whitelist=()):
        un = _i64(uni_nbr)
        us = np.ascontiguousarray(uni_stat, dtype=np.float64)
        up = np.ascontiguousarray(uni_p, dtype=np.float64)
        wl = _i64(list(whitelist))
        cap = max(len(un), 1)
        pn = np.zeros(cap, np.int64)
        ps = np.zeros(cap)
        pp = np.zeros(cap)
        nt = C.c_int64(0)
        k = self.L.fwo_hiton_pc(self.h, T, _ptr(un), _ptr(us), _ptr(up), len(un), max_k, alpha, hps, n_obs_min, max_tests,
                                _ptr(wl), len(wl), _ptr(pn), _ptr(ps), _ptr(pp), C.byref(nt))
        return pn[:k].copy(), ps[:k].copy(), pp[:k].copy(), int(nt.value)

    def hiton_pc_batch(self, targets, uni_off, uni_nbr, uni_stat, uni_p, max_k=3, alpha=0.01, hps=5, n_obs_min=0, max_tests=0, n_threads=0):
        """si_HITON_PC ("single" semantics) for many targets; CSR in, CSR out over the same offsets:
        returns (pc_count, pc_nbr, pc_stat, pc_p, num_tests)."""
        tg, off, un = _i64(targets), _i64(uni_off), _i64(uni_nbr)
        us = np.ascontiguousarray(uni_stat, dtype=np.float64)
        up = np.ascontiguousarray(uni_p, dtype=np.float64)
        nt, cap = len(tg), max(int(off[-1]), 1)
        pcc = np.zeros(max(nt, 1), np.int64); ntests = np.zeros(max(nt, 1), np.int64)
        pn = np.zeros(cap, np.int64); ps = np.zeros(cap); pp = np.zeros(cap)
        if n_threads <= 0:
            n_threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        self.L.fwo_hiton_pc_batch(self.h, nt, _ptr(tg), _ptr(off), _ptr(un), _ptr(us), _ptr(up), max_k, alpha, hps, n_obs_min, max_tests,
                                  n_threads, _ptr(pcc), _ptr(pn), _ptr(ps), _ptr(pp), _ptr(ntests))
        return pcc[:nt], pn, ps, pp, ntests[:nt]

    def lgl(self, max_k=3, alpha=0.01, hps=5, n_obs_min=-1, max_tests=10_000_000, fdr=True, mode="single", n_threads=1,
            targets=None, want_pc=False):
        p = self.p
        cap = p * (p - 1) // 2 if p < 4096 else 64 * p
        ea = np.zeros(cap, np.int64)
        eb = np.zeros(cap, np.int64)
        ew = np.zeros(cap)
        ct = C.c_int64(0)
        pt = C.c_int64(0)
        tg = None if targets is None else _i64(targets)
        pco = pcn = pcs = pcp = None
        max_pc = 0
        secs = np.zeros(3)
        if want_pc:
            max_pc = cap * 2
            pco = np.zeros(p + 1, np.int64)
            pcn = np.zeros(max_pc, np.int64)
            pcs = np.zeros(max_pc)
            pcp = np.zeros(max_pc)
        ne = self.L.fwo_lgl(self.h, max_k, alpha, hps, n_obs_min, max_tests, int(fdr), {"single": 0, "single_il": 1}[mode], n_threads,
                            _ptr(tg), 0 if tg is None else len(tg), _ptr(ea), _ptr(eb), _ptr(ew), cap, C.byref(ct), C.byref(pt),
                            _ptr(pco), _ptr(pcn), _ptr(pcs), _ptr(pcp), max_pc, _ptr(secs))
        assert ne >= 0, "edge capacity exceeded"
        res = {"edges": [(int(a), int(b), float(w)) for a, b, w in zip(ea[:ne], eb[:ne], ew[:ne])],
               "cond_tests": int(ct.value), "pair_tests": int(pt.value),
               "secs": {"cor": float(secs[0]), "pairwise": float(secs[1]), "hiton": float(secs[2])}}
        if want_pc:
            res["pc"] = (pco, pcn[:pco[-1]], pcs[:pco[-1]], pcp[:pco[-1]])
        return res


def fz_pval(stat, n, len_z=0):
    return lib().fwo_fz_pval(stat, n, len_z)


def chisq_sf(df, x):
    return lib().fwo_chisq_sf(df, x)


def mi_pval(mi, df, n_obs):
    return lib().fwo_mi_pval(mi, df, n_obs)


def benjamini_hochberg(pvals, alpha=0.01, m=None):
    pv = np.ascontiguousarray(pvals, dtype=np.float64).copy()
    lib().fwo_benjamini_hochberg(_ptr(pv), len(pv), alpha, len(pv) if m is None else m)
    return pv


def pcor_rec(cor, X, Y, Zs, cont32=True):
    c = np.asfortranarray(np.asarray(cor, dtype=np.float64))
    Zs = _i64(Zs)
    ns = C.c_int64(0)
    return lib().fwo_pcor_rec(_ptr(c), c.shape[0], int(cont32), X, Y, _ptr(Zs), len(Zs), C.byref(ns))


def mutual_information(ctab):
    t = np.asfortranarray(np.asarray(ctab, dtype=np.int64))
    if t.ndim == 2:
        return lib().fwo_mutual_information(_ptr(t), t.shape[0], t.shape[1], 0)
    return lib().fwo_mutual_information(_ptr(t), t.shape[0], t.shape[1], t.shape[2])


def num_threads():
    return lib().fwo_num_threads()
