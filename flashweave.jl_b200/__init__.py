"""flashweave.jl_b200 — host side of the B200 CI-test engine (ctypes over libfwgpu.so).

Mirrors the reference's operator interface for the hot path (names, argument meaning and
error behaviour), 0-based indices:

    reference (Julia, 1-based)                          here
    ---------------------------------------------------------------------------------
    test(X, Y, Zs, data, test_obj, ...)   tests.jl:28-265    Engine.test / Engine.test_batch
    test_subsets(X, Y, Z_total, ...)      tests.jl:281-346   Engine.test_subsets(_batch)
    pw_univar_neighbors(data; ...)        tests.jl:436-532   Engine.pw_univar_neighbors
    cor(data) -> Matrix{Float32}          learning.jl:42-44  Engine.cor
    si_HITON_PC(T, data, ...)             hiton.jl:283-400   Engine.si_HITON_PC
    LGL(data; parallel="single", ...)     learning.jl:203-279 Engine.LGL

There is no CPU fallback: every compute entry point goes through libfwgpu.so and raises
FwError when the library or a CUDA device is missing.  (The directory name contains a dot,
so import it through `fwload.load()` at the repo root, which registers it as
`flashweave_jl_b200`.)
"""
import ctypes as C
import math
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

KINDS = {"mi": 0, "mi_nz": 1, "fz": 2, "fz_nz": 3}
FW_OK = 0

# every symbol include/fwgpu.h declares (checked by tests/test_abi.py)
ABI_SYMBOLS = [
    "fw_create", "fw_destroy", "fw_last_error", "fw_set_index_base", "fw_stream", "fw_synchronize", "fw_launch_count", "fw_last_timing",
    "fw_hiton_exec_by_k",
    "fw_set_data_f32", "fw_set_data_i32", "fw_adopt_data_f32_device", "fw_set_n_obs", "fw_levels", "fw_cor_matrix",
    "fw_set_cor_f32", "fw_adopt_cor_device", "fw_cor_device_ptr", "fw_adopt_cor_device_rows", "fw_cor_prepare", "fw_cor_rows", "fw_cor_symmetrize", "fw_upload_cor_f32", "fw_test_batch", "fw_test_subsets", "fw_test_subsets_batch",
    "fw_pairwise", "fw_pairwise_copy", "fw_pairwise_partial", "fw_pairwise_partial_copy", "fw_pairwise_merge", "fw_set_univar_nbrs", "fw_pairwise_stats", "fw_hiton_pc", "fw_hiton_pc_ex", "fw_hiton_pc_capacity",
    "fw_normalize_f32", "fw_get_data_f32", "fw_get_data_i32", "fw_set_data_csc_f32", "fw_set_data_csc_i32",
    "fw_host_register", "fw_host_unregister",
    "fw_cor_gather", "fw_pairwise_prefetch", "fw_set_meta_mask", "fw_get_meta_mask", "fw_set_semantics",
    "fw_comm_handle_bytes", "fw_comm_export", "fw_comm_attach", "fw_comm_detach", "fw_multi_set_data_f32", "fw_multi_cor",
    "fw_build_info",
]

# normalize_data's mode names and per-test defaults (src/preprocessing.jl:666-668, 569-573) -> enum fw_norm_mode
NORM_MODES = {"tss": 0, "clr-adapt": 1, "clr-nonzero": 2, "pres-abs": 3, "clr-nonzero-binned": 4, "tss-nonzero-binned": 5}
DEFAULT_NORM = {"mi": "pres-abs", "mi_nz": "clr-nonzero-binned", "fz": "clr-adapt", "fz_nz": "clr-nonzero"}
NORM_KIND = {"tss": "fz", "clr-adapt": "fz", "clr-nonzero": "fz_nz", "pres-abs": "mi", "clr-nonzero-binned": "mi_nz", "tss-nonzero-binned": "mi_nz"}


class FwError(RuntimeError):
    pass


class TestResult(C.Structure):
    """src/types.jl:140-145"""
    _fields_ = [("stat", C.c_double), ("pval", C.c_double), ("df", C.c_int64),
                ("suff_power", C.c_uint8), ("_pad", C.c_uint8 * 7)]

    def astuple(self):
        return (self.stat, self.pval, int(self.df), bool(self.suff_power))

    def __repr__(self):
        return "TestResult(stat=%r, pval=%r, df=%d, suff_power=%s)" % self.astuple()


def lib_path():
    # FW_LIB_PATH: kernel-tuning experiments load an alternative build of the same library (scripts/); never a fallback
    return os.environ.get("FW_LIB_PATH") or os.path.join(_HERE, "libfwgpu.so")


def load_library():
    """dlopen libfwgpu.so (built in-tree by flashweave.jl_b200/build.py). Fails loudly."""
    global _LIB
    if _LIB is not None:
        return _LIB
    so = lib_path()
    if not os.path.exists(so):
        raise FwError("libfwgpu.so is not built (%s missing); run `python __graft_entry__.py` or flashweave.jl_b200/build.py. "
                      "There is no CPU fallback." % so)
    L = C.CDLL(so)
    i32, i64, dbl, vp = C.c_int32, C.c_int64, C.c_double, C.c_void_p
    sig = {
        "fw_create": (i32, [i32, C.POINTER(vp)]),
        "fw_destroy": (i32, [vp]),
        "fw_last_error": (C.c_char_p, [vp]),
        "fw_set_index_base": (i32, [vp, i32]),
        "fw_stream": (vp, [vp]),
        "fw_synchronize": (i32, [vp]),
        "fw_launch_count": (i64, [vp]),
        "fw_last_timing": (i32, [vp, vp, i32]),
        "fw_hiton_exec_by_k": (i32, [vp, vp]),
        "fw_set_data_f32": (i32, [vp, vp, i64, i64, i64]),
        "fw_set_data_i32": (i32, [vp, vp, i64, i64, i64]),
        "fw_adopt_data_f32_device": (i32, [vp, vp, i64, i64, i64]),
        "fw_set_n_obs": (i32, [vp, i64]),
        "fw_levels": (i32, [vp, vp, vp]),
        "fw_cor_matrix": (i32, [vp, vp]),
        "fw_set_cor_f32": (i32, [vp, vp, i64]),
        "fw_adopt_cor_device": (i32, [vp, vp, i64]),
        "fw_cor_device_ptr": (vp, [vp]),
        "fw_adopt_cor_device_rows": (i32, [vp, vp, i64, i64]),
        "fw_cor_prepare": (i32, [vp, vp]),
        "fw_cor_rows": (i32, [vp, i32, i32]),
        "fw_cor_symmetrize": (i32, [vp]),
        "fw_upload_cor_f32": (i32, [vp, vp, i64, i64, i64, vp]),
        "fw_test_batch": (i32, [vp, i32, i64, vp, vp, vp, vp, i64, i64, vp]),
        "fw_test_subsets": (i32, [vp, i32, i64, i64, vp, i64, i32, dbl, i64, i64, i64, vp, vp, vp, vp, vp]),
        "fw_test_subsets_batch": (i32, [vp, i32, i64, vp, vp, vp, vp, i32, dbl, i64, i64, i64, vp, vp, vp, vp, vp]),
        "fw_pairwise": (i32, [vp, i32, dbl, i64, i64, i32, i32, vp]),
        "fw_pairwise_copy": (i32, [vp, vp, vp, vp, vp]),
        "fw_set_univar_nbrs": (i32, [vp, vp, vp, vp, vp]),
        "fw_pairwise_stats": (i32, [vp, vp, vp, vp]),
        "fw_hiton_pc": (i32, [vp, i32, i64, vp, i32, dbl, i64, i64, i64] + [vp] * 11),
        "fw_hiton_pc_ex": (i32, [vp, i32, i64, vp, i32, dbl, i64, i64, i64] + [vp] * 4 + [vp] * 11 + [vp] * 7),
        "fw_hiton_pc_capacity": (i32, [vp, i64, vp, vp]),
        "fw_normalize_f32": (i32, [vp, vp, i64, i64, i64, i32, i32, C.POINTER(i64), C.POINTER(i64), vp, vp]),
        "fw_set_data_csc_f32": (i32, [vp, vp, vp, vp, i64, i64]),
        "fw_set_data_csc_i32": (i32, [vp, vp, vp, vp, i64, i64]),
        "fw_host_register": (i32, [vp, vp, i64]),
        "fw_host_unregister": (i32, [vp, vp]),
        "fw_get_data_f32": (i32, [vp, vp, i64]),
        "fw_get_data_i32": (i32, [vp, vp, i64]),
        "fw_cor_gather": (i32, [vp, vp, i64, vp]),
        "fw_set_semantics": (i32, [vp, i32]),
        "fw_set_meta_mask": (i32, [vp, vp, i64]),
        "fw_get_meta_mask": (i32, [vp, vp, i64]),
        "fw_pairwise_prefetch": (i32, [vp, dbl, i64]),
        "fw_pairwise_partial": (i32, [vp, i32, dbl, i64, i64, i32, i32, i32, vp, vp]),
        "fw_pairwise_partial_copy": (i32, [vp, vp, vp, vp, vp]),
        "fw_pairwise_merge": (i32, [vp, i32, dbl, i32, i64, vp, vp, vp, vp, i64, vp]),
        "fw_comm_handle_bytes": (i32, []),
        "fw_comm_export": (i32, [vp, i32, i32, i64, i64, vp]),
        "fw_comm_attach": (i32, [vp, vp]),
        "fw_comm_detach": (i32, [vp]),
        "fw_multi_set_data_f32": (i32, [vp, vp, i64]),
        "fw_multi_cor": (i32, [vp]),
        "fw_build_info": (C.c_char_p, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _LIB = L
    return L


def _p(a):
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


def _i64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int64))


class NbrCSR:
    """var -> OrderedDict(nbr -> (stat, adj p)) of pw_univar_neighbors (tests.jl:372-388) as CSR."""

    def __init__(self, offsets, nbr, stat, pval):
        self.offsets, self.nbr, self.stat, self.pval = offsets, nbr, stat, pval

    def __len__(self):
        return len(self.offsets) - 1

    def __getitem__(self, v):
        a, b = self.offsets[v], self.offsets[v + 1]
        return {int(n): (float(s), float(p)) for n, s, p in zip(self.nbr[a:b], self.stat[a:b], self.pval[a:b])}

    def degree(self):
        return np.diff(self.offsets)


class HitonResult:
    """Per-target HitonState.state_results / inter_results (types.jl:154-160) as CSR over the listed targets."""

    def __init__(self, targets, off, pc_count, pc_nbr, pc_stat, pc_p, tpc_count, tpc_nbr, tpc_stat, tpc_p, num_tests, executed):
        self.targets, self.off = targets, off
        self.pc_count, self.pc_nbr, self.pc_stat, self.pc_p = pc_count, pc_nbr, pc_stat, pc_p
        self.tpc_count, self.tpc_nbr, self.tpc_stat, self.tpc_p = tpc_count, tpc_nbr, tpc_stat, tpc_p
        self.num_tests, self.tests_executed = num_tests, executed

    def pc(self, i):
        a = self.off[i]
        b = a + self.pc_count[i]
        return self.pc_nbr[a:b], self.pc_stat[a:b], self.pc_p[a:b]

    def tpc(self, i):
        a = self.off[i]
        b = a + self.tpc_count[i]
        return self.tpc_nbr[a:b], self.tpc_stat[a:b], self.tpc_p[a:b]

    def rejections(self, i):
        """HitonState.state_rejections of target i (track_rejections=True): {candidate: (Zs, TestResult tuple, (num_tests, frac))}"""
        r = self.rej
        a = self.off[i]
        out = {}
        for j in range(a, a + int(r["count"][i])):
            out[int(r["nbr"][j])] = (tuple(int(z) for z in r["Zs"][j, :r["k"][j]]), r["res"][j].astuple(), (int(r["ntests"][j]), float(r["frac"][j])))
        return out


class Engine:
    """One `fw_ctx`: the reference's test_obj + data + cor_mat, resident on one B200."""

    def __init__(self, device=0):
        self.L = load_library()
        h = C.c_void_p()
        st = self.L.fw_create(device, C.byref(h))
        if st != FW_OK:
            raise FwError("fw_create(device=%d) failed [%d]: %s" % (device, st, self.L.fw_last_error(None).decode()))
        self.h = h
        self.device = device
        self.kind = None
        self.n = self.p = 0
        self._keep = []

    def _unpin_hbuf(self):
        b = getattr(self, "_hbuf", None)
        if b is not None and getattr(self, "h", None):
            for k in b.get("pinned", []):
                self.L.fw_host_unregister(self.h, _p(b[k]))
            b["pinned"] = []

    def close(self):
        self._unpin_hbuf()
        if getattr(self, "h", None):
            self.L.fw_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, st):
        if st != FW_OK:
            raise FwError("libfwgpu error [%d]: %s" % (st, self.L.fw_last_error(self.h).decode()))

    # -- plumbing -------------------------------------------------------------------------
    @property
    def stream(self):
        return self.L.fw_stream(self.h)

    def synchronize(self):
        self._ck(self.L.fw_synchronize(self.h))

    def launch_count(self):
        return int(self.L.fw_launch_count(self.h))

    def last_timing(self):
        """device ms of the last cor / pairwise / hiton phases (CUDA events on the engine's stream)"""
        out = np.zeros(6)
        self._ck(self.L.fw_last_timing(self.h, _p(out), 6))
        return {"cor_ms": out[0], "pairwise_ms": out[1], "hiton_ms": out[2], "cor_standardise_ms": out[4], "cor_barrier_wait_ms": out[5]}

    def hiton_exec_by_k(self):
        out = np.zeros(3, np.int64)
        self._ck(self.L.fw_hiton_exec_by_k(self.h, _p(out)))
        return out

    # -- data -----------------------------------------------------------------------------
    def set_data(self, data, kind):
        """data: [n, p] array (any layout) or an already column-major buffer given as a [p, n] C-contiguous
        array with `data.T` semantics via set_data_colmajor."""
        d = np.asarray(data)
        return self.set_data_colmajor(np.ascontiguousarray(d.T), kind)

    def set_data_colmajor(self, data_pn, kind):
        """data_pn: C-contiguous [p, n] (= Julia's column-major n x p Matrix)."""
        k = KINDS[kind]
        p, n = data_pn.shape
        if k >= 2:
            if data_pn.dtype != np.float32 or not data_pn.flags.c_contiguous:
                data_pn = np.ascontiguousarray(data_pn, dtype=np.float32)
            self._ck(self.L.fw_set_data_f32(self.h, _p(data_pn), n, p, n))
        else:
            if data_pn.dtype != np.int32 or not data_pn.flags.c_contiguous:
                data_pn = np.ascontiguousarray(data_pn, dtype=np.int32)
            self._ck(self.L.fw_set_data_i32(self.h, _p(data_pn), n, p, n))
        self._keep = [data_pn]
        self.kind, self.n, self.p = kind, n, p
        self._cor_valid = False
        return self

    def normalize_data(self, data, test_name="", norm_mode="", n_bins=3, want_host=True, meta_data=None, meta_header=None, make_onehot=True):
        """normalize_data(data; test_name | norm_mode) of the reference (src/preprocessing.jl:660-684) for a dense [n, p] table of
        counts without meta variables.  The normalised table stays resident as the engine's table (ready for cor / pairwise /
        HITON-PC with the matching test kind); returns dict(data=[n', p'] array or None, col_mask, obs_filter_mask, kind)."""
        if bool(test_name) == bool(norm_mode):
            raise ValueError("provide either test_name or norm_mode (but not both)")
        mode = norm_mode or DEFAULT_NORM[test_name]
        if mode not in NORM_MODES:
            raise ValueError("%s is not a valid normalization mode" % mode)
        d = np.ascontiguousarray(np.asarray(data).T, dtype=np.float32)          # [p, n] = column-major n x p Matrix{Float32}
        p, n = d.shape
        rmask = np.zeros(n, np.uint8); cmask = np.zeros(p, np.uint8)
        n_out, p_out = C.c_int64(0), C.c_int64(0)
        self._ck(self.L.fw_normalize_f32(self.h, _p(d), n, p, n, NORM_MODES[mode], n_bins, C.byref(n_out), C.byref(p_out), _p(rmask), _p(cmask)))
        self.kind = test_name or NORM_KIND[mode]
        self.n, self.p = n_out.value, p_out.value
        self._cor_valid = False
        out = None
        if want_host and self.n > 0 and self.p > 0:
            if NORM_MODES[mode] <= 2:
                out = np.empty((self.p, self.n), np.float32)
                self._ck(self.L.fw_get_data_f32(self.h, _p(out), self.n))
            else:
                out = np.empty((self.p, self.n), np.int32)
                self._ck(self.L.fw_get_data_i32(self.h, _p(out), self.n))
            out = out.T
        res = {"data": out, "col_mask": cmask.astype(bool), "obs_filter_mask": rmask.astype(bool), "kind": self.kind,
               "meta_mask": np.zeros(self.p, bool), "meta_names": []}
        if meta_data is not None and self.n > 0 and self.p > 0:
            # meta variables (preprocessing.jl:418-446, 523-556) are host-side factor handling: the normalised OTU table comes back
            # once, the prepared meta columns are appended, and the combined table becomes the resident one
            from . import meta as _meta
            table = out if out is not None else self.get_data()
            comb, mmask, mnames = _meta.combine_with_meta(table, res["obs_filter_mask"], meta_data, meta_header, mode, make_onehot)
            self.set_data(comb, self.kind)
            self.set_meta_mask(mmask)
            res.update(data=comb if want_host else None, meta_mask=mmask, meta_names=mnames)
        return res

    def set_data_csc(self, colptr, rowval, nzval, n, p, kind):
        """SparseMatrixCSC{T,Int64} triple (0-based here; the Julia glue sets index base 1) of an n x p table."""
        cp, rv = _i64(colptr), _i64(rowval)
        if KINDS[kind] >= 2:
            nz = np.ascontiguousarray(nzval, dtype=np.float32)
            self._ck(self.L.fw_set_data_csc_f32(self.h, _p(cp), _p(rv), _p(nz), n, p))
        else:
            nz = np.ascontiguousarray(nzval, dtype=np.int32)
            self._ck(self.L.fw_set_data_csc_i32(self.h, _p(cp), _p(rv), _p(nz), n, p))
        self.kind, self.n, self.p = kind, n, p
        self._cor_valid = False
        return self

    def get_data(self):
        """the resident table as an [n, p] array"""
        if KINDS[self.kind] >= 2:
            out = np.empty((self.p, self.n), np.float32); self._ck(self.L.fw_get_data_f32(self.h, _p(out), self.n))
        else:
            out = np.empty((self.p, self.n), np.int32); self._ck(self.L.fw_get_data_i32(self.h, _p(out), self.n))
        return out.T

    def set_data_ptr(self, host_ptr, n, p, kind="fz"):
        """host pointer (e.g. a pinned torch tensor's data_ptr) to a column-major n x p float32 table"""
        assert KINDS[kind] >= 2
        self._ck(self.L.fw_set_data_f32(self.h, C.c_void_p(host_ptr), n, p, n))
        self.kind, self.n, self.p = kind, n, p
        self._cor_valid = False
        return self

    def adopt_data_device(self, dev_ptr, n, p, kind="fz"):
        self._ck(self.L.fw_adopt_data_f32_device(self.h, C.c_void_p(dev_ptr), n, p, n))
        self.kind, self.n, self.p = kind, n, p
        self._cor_valid = False
        return self

    def set_semantics(self, sparse):
        """mi_nz: follow the reference's sparse-input code path (contingency.jl:182-258, 300-480) instead of the dense one"""
        self._ck(self.L.fw_set_semantics(self.h, 1 if sparse else 0))

    def set_meta_mask(self, mask):
        m = np.ascontiguousarray(np.asarray(mask, dtype=bool).astype(np.uint8))
        self._ck(self.L.fw_set_meta_mask(self.h, _p(m), len(m)))

    def meta_mask(self):
        m = np.zeros(self.p, np.uint8)
        self._ck(self.L.fw_get_meta_mask(self.h, _p(m), self.p))
        return m.astype(bool)

    def levels(self):
        """get_levels / get_max_vals (misc.jl:64-97) of the resident discrete table"""
        lv = np.zeros(self.p, np.int32)
        mv = np.zeros(self.p, np.int32)
        self._ck(self.L.fw_levels(self.h, _p(lv), _p(mv)))
        return lv, mv

    def upload_and_cor(self, host_ptr_or_array, n=None, p=None, want_host=False):
        """set_data + cor in one call with the upload hidden behind the GEMM (fz).  Accepts a C-contiguous [p, n] float32 array
        or a raw host pointer (pinned memory gives full PCIe speed) together with n and p."""
        if isinstance(host_ptr_or_array, np.ndarray):
            a = host_ptr_or_array
            assert a.dtype == np.float32 and a.flags.c_contiguous
            p, n = a.shape
            ptr = a.ctypes.data
            self._keep = [a]
        else:
            ptr = int(host_ptr_or_array)
        out = np.empty((p, p), np.float32) if want_host else None
        self._ck(self.L.fw_upload_cor_f32(self.h, C.c_void_p(ptr), n, p, n, _p(out)))
        self.kind, self.n, self.p = "fz", n, p
        self._cor_valid = True
        return out

    def set_n_obs(self, n):
        self._ck(self.L.fw_set_n_obs(self.h, n))
        self.n = n

    # -- cor_mat ----------------------------------------------------------------------------
    def cor(self, want_host=True):
        """cor_mat = Float32.(cor(data)) (learning.jl:42-44), computed and kept on the device."""
        out = np.empty((self.p, self.p), np.float32) if want_host else None
        self._ck(self.L.fw_cor_matrix(self.h, _p(out)))
        self._cor_valid = True
        return out

    def set_cor(self, cor, n_obs=None, kind="fz"):
        c = np.ascontiguousarray(cor, dtype=np.float32)
        assert c.ndim == 2 and c.shape[0] == c.shape[1]
        self._ck(self.L.fw_set_cor_f32(self.h, _p(c), c.shape[0]))
        self.p = c.shape[0]
        self._cor_valid = True
        if self.kind is None:
            self.kind = kind
        if n_obs is not None:
            self.set_n_obs(n_obs)
        self.synchronize()
        return self

    def adopt_cor_device(self, dev_ptr, p):
        self._ck(self.L.fw_adopt_cor_device(self.h, C.c_void_p(dev_ptr), p))
        self.p = p
        self._cor_valid = True

    def cor_device_ptr(self):
        return self.L.fw_cor_device_ptr(self.h)

    def cor_gather(self, idx):
        """cor_mat[idx][:, idx] as an [m, m] float32 array (works on the full and on the row-sharded matrix)"""
        ix = _i64(idx)
        out = np.empty((len(ix), len(ix)), np.float32)
        self._ck(self.L.fw_cor_gather(self.h, _p(ix), len(ix), _p(out)))
        return out

    def pairwise_prefetch(self, alpha=0.01, n_obs_min=0):
        """announce the next pw_univar_neighbors(fz): the cor_mat GEMM collects its raw candidates in the epilogue"""
        self._ck(self.L.fw_pairwise_prefetch(self.h, alpha, n_obs_min))

    # -- multi-GPU group (include/fwgpu.h; parallel.attach_group does the handle exchange) -----------------------
    def comm_export(self, rank, world, n, p):
        buf = np.zeros(self.L.fw_comm_handle_bytes(), np.uint8)
        self._ck(self.L.fw_comm_export(self.h, rank, world, n, p, _p(buf)))
        return buf

    def comm_attach(self, handles):
        hb = np.ascontiguousarray(handles, dtype=np.uint8)
        self._ck(self.L.fw_comm_attach(self.h, _p(hb)))

    def comm_detach(self):
        self._ck(self.L.fw_comm_detach(self.h))

    def multi_set_data_ptr(self, host_slice_ptr, n, p, ld=None):
        """this rank's columns of the table (host pointer at column p*rank/world, column-major, leading dimension ld)"""
        self._ck(self.L.fw_multi_set_data_f32(self.h, C.c_void_p(host_slice_ptr), ld or n))
        self.kind, self.n, self.p = "fz", n, p
        self._cor_valid = False

    def multi_cor(self):
        self._ck(self.L.fw_multi_cor(self.h))
        self._cor_valid = True

    # row-sharded cor_mat with a host-side exchange (NCCL all-gather by the caller): see parallel.sharded_cor
    def adopt_cor_device_rows(self, dev_ptr, p, rows_allocated):
        self._ck(self.L.fw_adopt_cor_device_rows(self.h, C.c_void_p(dev_ptr), p, rows_allocated))
        self.p = p
        self._cor_valid = True

    def cor_prepare(self):
        nb = C.c_int32(0)
        self._ck(self.L.fw_cor_prepare(self.h, C.byref(nb)))
        return int(nb.value)

    def cor_rows(self, tile_row_begin, tile_row_end):
        self._ck(self.L.fw_cor_rows(self.h, tile_row_begin, tile_row_end))

    def cor_symmetrize(self):
        self._ck(self.L.fw_cor_symmetrize(self.h))

    # -- tests ------------------------------------------------------------------------------
    def test_batch(self, X, Y, Zs=None, k=None, hps=5, n_obs_min=0, kind=None):
        X, Y = _i64(X), _i64(Y)
        nt = len(X)
        if Zs is None:
            Zs = np.zeros((nt, 3), np.int64)
            k = np.zeros(nt, np.int32)
        else:
            Zl = list(Zs)
            if k is None:
                k = np.array([len(z) for z in Zl], np.int32)
            Zs = np.zeros((nt, 3), np.int64)
            for i, z in enumerate(Zl):
                Zs[i, :len(z)] = z
        k = np.ascontiguousarray(k, dtype=np.int32)
        out = (TestResult * max(nt, 1))()
        self._ck(self.L.fw_test_batch(self.h, KINDS[kind or self.kind], nt, _p(X), _p(Y), _p(k), _p(np.ascontiguousarray(Zs)), hps, n_obs_min, out))
        return [out[i].astuple() for i in range(nt)]

    def test(self, X, Y, Zs=(), hps=5, n_obs_min=0, kind=None):
        return self.test_batch([X], [Y], [tuple(Zs)], hps=hps, n_obs_min=n_obs_min, kind=kind)[0]

    def test_subsets(self, X, Y, Z_total, max_k=3, alpha=0.01, hps=5, n_obs_min=0, max_tests=0, kind=None):
        Z = _i64(Z_total)
        out = TestResult()
        Zs = np.zeros(3, np.int64)
        k = C.c_int32(0)
        nt = C.c_int64(0)
        fr = C.c_double(0)
        self._ck(self.L.fw_test_subsets(self.h, KINDS[kind or self.kind], X, Y, _p(Z), len(Z), max_k, alpha, hps, n_obs_min, max_tests,
                                        C.byref(out), _p(Zs), C.byref(k), C.byref(nt), C.byref(fr)))
        return out.astuple(), tuple(int(z) for z in Zs[:k.value]), int(nt.value), fr.value

    def test_subsets_batch(self, X, Y, Z_lists, max_k=3, alpha=0.01, hps=5, n_obs_min=0, max_tests=0, kind=None):
        X, Y = _i64(X), _i64(Y)
        nj = len(X)
        off = np.zeros(nj + 1, np.int64)
        off[1:] = np.cumsum([len(z) for z in Z_lists])
        zi = _i64(np.concatenate([np.asarray(z, np.int64) for z in Z_lists]) if off[-1] else np.zeros(0, np.int64))
        out = (TestResult * max(nj, 1))()
        Zs = np.zeros((max(nj, 1), 3), np.int64)
        k = np.zeros(max(nj, 1), np.int32)
        nt = np.zeros(max(nj, 1), np.int64)
        fr = np.zeros(max(nj, 1), np.float64)
        self._ck(self.L.fw_test_subsets_batch(self.h, KINDS[kind or self.kind], nj, _p(X), _p(Y), _p(off), _p(zi), max_k, alpha, hps,
                                              n_obs_min, max_tests, out, _p(Zs), _p(k), _p(nt), _p(fr)))
        return [(out[i].astuple(), tuple(int(z) for z in Zs[i, :k[i]]), int(nt[i]), float(fr[i])) for i in range(nj)]

    # -- pairwise stage ------------------------------------------------------------------------
    def pw_univar_neighbors(self, alpha=0.01, hps=5, n_obs_min=0, FDR=True, correct_reliable_only=True, want_host=True, kind=None):
        ne = C.c_int64(0)
        self._ck(self.L.fw_pairwise(self.h, KINDS[kind or self.kind], alpha, hps, n_obs_min, int(FDR), int(correct_reliable_only), C.byref(ne)))
        self.uni_entries = int(ne.value)
        if not want_host:
            return None
        return self.univar_nbrs()

    def pairwise_partial(self, rank, world, alpha=0.01, hps=5, n_obs_min=0, correct_reliable_only=True, kind=None):
        """This rank's share of pw_univar_neighbors for the table-based kinds (fw_pairwise_partial): the raw-significant records
        {"x", "y", "p", "stat"} of the pairs whose X is dealt to `rank`, and the number of its tests that enter the correction."""
        nr, nrel = C.c_int64(0), C.c_int64(0)
        self._ck(self.L.fw_pairwise_partial(self.h, KINDS[kind or self.kind], alpha, hps, n_obs_min, int(correct_reliable_only), rank, world,
                                            C.byref(nr), C.byref(nrel)))
        n = int(nr.value)
        rec = {"x": np.zeros(n, np.int32), "y": np.zeros(n, np.int32), "p": np.zeros(n), "stat": np.zeros(n), "n_reliable": int(nrel.value)}
        self._ck(self.L.fw_pairwise_partial_copy(self.h, _p(rec["x"]), _p(rec["y"]), _p(rec["p"]), _p(rec["stat"])))
        return rec

    def pairwise_partial_run(self, rank, world, alpha=0.01, hps=5, n_obs_min=0, correct_reliable_only=True, kind=None):
        """fw_pairwise_partial without fetching the records: (n_raw, n_reliable); the records stay on the device"""
        nr, nrel = C.c_int64(0), C.c_int64(0)
        self._ck(self.L.fw_pairwise_partial(self.h, KINDS[kind or self.kind], alpha, hps, n_obs_min, int(correct_reliable_only), rank, world,
                                            C.byref(nr), C.byref(nrel)))
        return int(nr.value), int(nrel.value)

    def pairwise_partial_copy_ptrs(self, x_ptr, y_ptr, p_ptr, stat_ptr):
        """this rank's records into caller-provided int32 / int32 / float64 / float64 buffers (host or device pointers)"""
        self._ck(self.L.fw_pairwise_partial_copy(self.h, C.c_void_p(x_ptr), C.c_void_p(y_ptr), C.c_void_p(p_ptr), C.c_void_p(stat_ptr)))

    def pairwise_merge_ptrs(self, n_total, x_ptr, y_ptr, p_ptr, stat_ptr, m_tests, alpha=0.01, FDR=True, kind=None):
        """fw_pairwise_merge on raw pointers (host or device memory)"""
        ne = C.c_int64(0)
        self._ck(self.L.fw_pairwise_merge(self.h, KINDS[kind or self.kind], alpha, int(FDR), n_total, C.c_void_p(x_ptr), C.c_void_p(y_ptr),
                                          C.c_void_p(p_ptr), C.c_void_p(stat_ptr), m_tests, C.byref(ne)))
        self.uni_entries = int(ne.value)

    def pairwise_merge(self, records, alpha=0.01, FDR=True, correct_reliable_only=True, kind=None, want_host=True):
        """BH + neighbour lists from the records of ALL ranks (fw_pairwise_merge); `records` = the dicts of pairwise_partial."""
        x = np.ascontiguousarray(np.concatenate([r["x"] for r in records]), dtype=np.int32)
        y = np.ascontiguousarray(np.concatenate([r["y"] for r in records]), dtype=np.int32)
        pv = np.ascontiguousarray(np.concatenate([r["p"] for r in records]), dtype=np.float64)
        st = np.ascontiguousarray(np.concatenate([r["stat"] for r in records]), dtype=np.float64)
        m = sum(int(r["n_reliable"]) for r in records) if correct_reliable_only else self.p * (self.p - 1) // 2
        ne = C.c_int64(0)
        self._ck(self.L.fw_pairwise_merge(self.h, KINDS[kind or self.kind], alpha, int(FDR), len(x), _p(x), _p(y), _p(pv), _p(st), m, C.byref(ne)))
        self.uni_entries = int(ne.value)
        return self.univar_nbrs() if want_host else None

    def univar_nbrs(self):
        ne = self.uni_entries
        off = np.zeros(self.p + 1, np.int64)
        nbr = np.zeros(max(ne, 1), np.int64)
        st = np.zeros(max(ne, 1))
        ap = np.zeros(max(ne, 1))
        self._ck(self.L.fw_pairwise_copy(self.h, _p(off), _p(nbr), _p(st), _p(ap)))
        return NbrCSR(off, nbr[:ne], st[:ne], ap[:ne])

    def set_univar_nbrs(self, offsets, nbr, stat, pval):
        off = _i64(offsets)
        self._ck(self.L.fw_set_univar_nbrs(self.h, _p(off), _p(_i64(nbr)), _p(np.ascontiguousarray(stat, dtype=np.float64)),
                                           _p(np.ascontiguousarray(pval, dtype=np.float64))))
        self.uni_entries = int(off[-1])

    def pairwise_stats(self):
        a, b, c = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        self._ck(self.L.fw_pairwise_stats(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return {"n_tests": a.value, "n_reliable": b.value, "n_raw_sig": c.value}

    # -- HITON-PC ---------------------------------------------------------------------------------
    def si_HITON_PC(self, targets, max_k=3, alpha=0.01, hps=5, n_obs_min=0, max_tests=10_000_000, kind=None, want_tpc=True,
                    buffers=None, reuse_buffers=False, whitelists=None, blacklists=None, track_rejections=False):
        """si_HITON_PC for each target (hiton.jl:283-400).  whitelists / blacklists: one iterable of variables per target
        (hiton.jl:20-38; empty everywhere = parallel="single" semantics); track_rejections: HitonResult.rejections(i)."""
        t = _i64(np.atleast_1d(targets))
        nt = len(t)
        cap = C.c_int64(0)
        self._ck(self.L.fw_hiton_pc_capacity(self.h, nt, _p(t), C.byref(cap)))
        cp = max(int(cap.value), 1)
        if buffers is None and reuse_buffers:
            buffers = getattr(self, "_hbuf", None)        # grow-only host buffers: the returned views alias them until the next call
        if buffers is not None and buffers["cap"] >= cp and buffers["nt"] >= nt:
            b = buffers
        else:
            b = {"cap": cp, "nt": nt, "off": np.zeros(nt + 1, np.int64), "pcc": np.zeros(max(nt, 1), np.int64),
                 "pcn": np.zeros(cp, np.int64), "pcs": np.zeros(cp), "pcp": np.zeros(cp),
                 "tpcc": np.zeros(max(nt, 1), np.int64), "tpcn": np.zeros(cp, np.int64), "tpcs": np.zeros(cp), "tpcp": np.zeros(cp),
                 "ntests": np.zeros(max(nt, 1), np.int64)}
            if reuse_buffers:
                self._unpin_hbuf()
                self._hbuf = b
                # reused result buffers are page-locked once: the copy-out then runs at full PCIe speed
                b["pinned"] = []
                for k in ("off", "pcc", "pcn", "pcs", "pcp", "tpcc", "tpcn", "tpcs", "tpcp", "ntests"):
                    if self.L.fw_host_register(self.h, _p(b[k]), b[k].nbytes) == 0:
                        b["pinned"].append(k)
        ex = C.c_int64(0)
        tp = want_tpc

        def csr(lists):
            if lists is None:
                return None, None
            assert len(lists) == nt, "one list per target"
            off = np.zeros(nt + 1, np.int64)
            off[1:] = np.cumsum([len(l) for l in lists])
            idx = _i64(np.concatenate([np.asarray(list(l), np.int64) for l in lists]) if off[-1] else np.zeros(0, np.int64))
            return off, idx
        wlo, wli = csr(whitelists)
        blo, bli = csr(blacklists)
        rej = None
        if track_rejections:
            rej = {"count": np.zeros(max(nt, 1), np.int64), "nbr": np.zeros(cp, np.int64), "Zs": np.zeros((cp, 3), np.int64), "k": np.zeros(cp, np.int32),
                   "res": (TestResult * cp)(), "ntests": np.zeros(cp, np.int64), "frac": np.zeros(cp)}
        self._ck(self.L.fw_hiton_pc_ex(self.h, KINDS[kind or self.kind], nt, _p(t), max_k, alpha, hps, n_obs_min, max_tests,
                                       _p(wlo), _p(wli), _p(blo), _p(bli),
                                       _p(b["off"]), _p(b["pcc"]), _p(b["pcn"]), _p(b["pcs"]), _p(b["pcp"]),
                                       _p(b["tpcc"]), _p(b["tpcn"]) if tp else None, _p(b["tpcs"]) if tp else None, _p(b["tpcp"]) if tp else None,
                                       _p(b["ntests"]), C.byref(ex),
                                       _p(rej["count"]) if rej else None, _p(rej["nbr"]) if rej else None, _p(rej["Zs"]) if rej else None,
                                       _p(rej["k"]) if rej else None, C.cast(rej["res"], C.c_void_p) if rej else None,
                                       _p(rej["ntests"]) if rej else None, _p(rej["frac"]) if rej else None))
        res = HitonResult(t, b["off"][:nt + 1], b["pcc"][:nt], b["pcn"], b["pcs"], b["pcp"], b["tpcc"][:nt], b["tpcn"], b["tpcs"], b["tpcp"],
                          b["ntests"][:nt], int(ex.value))
        res.rej = rej
        return res

    # -- LGL ------------------------------------------------------------------------------------------
    def LGL(self, max_k=3, alpha=0.01, hps=5, n_obs_min=-1, max_tests=10_000_000, FDR=True, targets=None, kind=None, parallel="single",
            track_rejections=False):
        """learning.jl:203-279: cor -> pairwise -> HITON-PC per target -> OR-rule graph.
        parallel="single" (learning.jl:137-138): targets are independent, one launch for all of them.
        parallel="single_il": the reference's default schedule with ONE worker (interleaved.jl:60-179): two initial jobs with empty
        whitelists, then one target at a time whose whitelist is its neighbourhood in the graph of all finished targets
        (feed-forward, interleaved.jl:124-128) - each job is one fw_hiton_pc_ex call, the loop is the host's."""
        kind = kind or self.kind
        if n_obs_min < 0:
            ml = int(self.levels()[0].max()) if kind in ("mi", "mi_nz") else None
            n_obs_min = auto_n_obs_min(kind, max_k, hps, max_level=ml)
        if kind == "fz" and not (getattr(self, "_cor_valid", False) and self.L.fw_cor_device_ptr(self.h)):
            self.cor(want_host=False)                      # no cor_mat yet, or it belongs to a previous table
        uni = self.pw_univar_neighbors(alpha=alpha, hps=hps, n_obs_min=n_obs_min, FDR=FDR, kind=kind)
        tg = target_order(uni) if targets is None else _i64(targets)
        kw = dict(max_k=max_k, alpha=alpha, hps=hps, n_obs_min=n_obs_min, max_tests=max_tests, kind=kind, want_tpc=False, track_rejections=track_rejections)
        if parallel == "single" or max_k == 0:
            res = self.si_HITON_PC(tg, **kw)
        elif parallel == "single_il":
            sched = list(tg[:2][::-1]) + list(tg[2:])      # the FIFO of the two initial jobs is served second-first (interleaved.jl:136-141)
            graph = {int(t): set() for t in range(self.p)}
            parts = []
            for i, T in enumerate(sched):
                wl = sorted(graph[int(T)]) if i >= 2 else []
                r = self.si_HITON_PC([T], whitelists=[wl], **kw)
                nb, st, pv = r.pc(0)
                parts.append((int(T), nb.copy(), st.copy(), pv.copy(), int(r.num_tests[0]), r.tests_executed, r.rejections(0) if track_rejections else None))
                for v in nb:
                    graph[int(T)].add(int(v)); graph[int(v)].add(int(T))
            res = _ListResult(parts)
        else:
            raise ValueError("parallel must be 'single' or 'single_il'")
        edges = assemble_graph(res, uni, kind)
        return {"edges": edges, "cond_tests": int(res.num_tests.sum()), "tests_executed": res.tests_executed,
                "pair_tests": self.p * (self.p - 1) // 2, "hiton": res, "univar": uni}


class _ListResult:
    """per-target results of sequential single-target calls, with the accessors of HitonResult"""

    def __init__(self, parts):
        self.targets = np.asarray([q[0] for q in parts], np.int64)
        self._pc = [(q[1], q[2], q[3]) for q in parts]
        self.num_tests = np.asarray([q[4] for q in parts], np.int64)
        self.tests_executed = int(sum(q[5] for q in parts))
        self._rej = [q[6] for q in parts]
        self.pc_count = np.asarray([len(q[1]) for q in parts], np.int64)

    def pc(self, i):
        return self._pc[i]

    def rejections(self, i):
        return self._rej[i]


# ---- host logic shared with the tests (pure Python, no compute) --------------------------------------
def auto_n_obs_min(kind, max_k, hps=5, max_level=None):
    """learning.jl:51-61 (applies to every test kind because of the `<` / `&` precedence quirk)."""
    if kind in ("mi", "mi_nz"):
        assert max_level is not None
        return hps * 2 * 2 * int(min(max_level ** max_k, 8))
    return 20


def target_order(uni):
    """learning.jl:97-98: variables by ascending univariate degree, stable."""
    return np.argsort(uni.degree(), kind="stable").astype(np.int64)


def shard_targets(order, rank, world):
    """Static interleaved sharding of the degree-ordered targets (target i -> rank i mod world);
    replaces the job queue of interleaved.jl:76-93 for parallel="single" semantics."""
    return np.ascontiguousarray(order[rank::world])


def _maxweight(w1, w2):
    # misc.jl:201-218
    if math.isnan(w1):
        return w2
    if math.isnan(w2):
        return w1
    s1 = (w1 > 0) - (w1 < 0)
    s2 = (w2 > 0) - (w2 < 0)
    if s1 * s2 < 0:
        return w1
    return max(abs(w1), abs(w2)) * s1


def assemble_graph(res, uni, kind):
    """make_weights (misc.jl:137-159) + make_symmetric_graph (misc.jl:230-272), OR rule.
    Returns sorted [(a, b, weight)] with a < b."""
    W = {}
    disc = kind in ("mi", "mi_nz")
    for i, T in enumerate(res.targets):
        nb, st, _ = res.pc(i)
        T = int(T)
        if disc:
            u = uni[T]
            W[T] = {int(v): float(np.sign(u[int(v)][0]) * abs(s)) for v, s in zip(nb, st)}
        else:
            W[T] = {int(v): float(s) for v, s in zip(nb, st)}
    edges = {}
    for a in sorted(W):
        for b, w in W[a].items():
            e = (min(a, b), max(a, b))
            if e in edges:
                continue
            rw = W.get(b, {}).get(a, float("nan"))
            sw = _maxweight(w, rw)
            if not math.isnan(sw):
                edges[e] = sw
    return sorted((a, b, w) for (a, b), w in edges.items())


def julia_float_str(x):
    """string(::Float64) of Julia (what src/io.jl:355 writes): shortest round-trip digits; positional notation for
    1e-4 <= |x| < 1e6 with at least one fractional digit, otherwise d.ddde[-]x (Base.Ryu.writeshortest as `show` calls it)"""
    x = float(x)
    if math.isnan(x):
        return "NaN"
    if math.isinf(x):
        return "Inf" if x > 0 else "-Inf"
    if x == 0.0:
        return "-0.0" if math.copysign(1.0, x) < 0 else "0.0"
    sign = "-" if x < 0 else ""
    digits, exp = ("%r" % abs(x)), 0
    m, _, e = digits.partition("e")
    exp = int(e) if e else 0
    ip, _, fp = m.partition(".")
    ds = (ip + fp).lstrip("0")
    point = len(ip) + exp - (len(ip + fp) - len((ip + fp).lstrip("0")))      # decimal exponent: value = 0.ds * 10^point
    ds = ds.rstrip("0") or "0"
    e10 = point - 1                                                             # value = d.ddd * 10^e10
    if -4 <= e10 < 6:
        if point <= 0:
            return sign + "0." + "0" * (-point) + ds
        if point >= len(ds):
            return sign + ds + "0" * (point - len(ds)) + ".0"
        return sign + ds[:point] + "." + ds[point:]
    return sign + ds[0] + "." + (ds[1:] or "0") + "e" + str(e10)


def write_edgelist(path, edges, header=None, meta_mask=None, p=None):
    """io.jl:338-359 edgelist format (`# header`, `# meta mask`, then `a<TAB>b<TAB>weight`).  Edges in the order `edges(G)` of the
    reference's SimpleWeightedGraph yields them: the upper triangle of the sparse weight matrix column by column, i.e. sorted by
    (larger endpoint, smaller endpoint)."""
    if header is None:
        header = ["X%d" % (i + 1) for i in range(p)]
    if meta_mask is None:
        meta_mask = [False] * len(header)
    with open(path, "w") as f:
        f.write("# header\t" + ",".join(header) + "\n")
        f.write("# meta mask\t" + ",".join("true" if m else "false" for m in meta_mask) + "\n")
        for a, b, w in sorted(((min(a, b), max(a, b), w) for a, b, w in edges), key=lambda e: (e[1], e[0])):
            f.write("%s\t%s\t%s\n" % (header[a], header[b], julia_float_str(w)))
