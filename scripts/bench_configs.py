"""Measures the BASELINE.json parity/size configs C1, C2, C3 and a C5-shaped fz_nz table on one GPU (not the bench.py
contract; fills BASELINE.md §5).  usage: python scripts/bench_configs.py [C1 C2 C3 C5s]"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fwload
fw = fwload.load(); synth = fwload.load_sub("synth")

def run(name, x, kind, max_k, reps=3):
    eng = fw.Engine(0)
    out = {"config": name, "kind": kind, "p": x.shape[0], "n": x.shape[1], "max_k": max_k}
    best = None
    for rep in range(reps):
        t0 = time.perf_counter(); eng.set_data_colmajor(x, kind); eng.synchronize(); t1 = time.perf_counter()
        if kind == "fz":
            eng.cor(want_host=False); eng.synchronize()
        t2 = time.perf_counter()
        nom = 20
        if kind in ("mi", "mi_nz"):
            nom = fw.auto_n_obs_min(kind, max_k, 5, max_level=int(eng.levels()[0].max()))
        eng.pw_univar_neighbors(alpha=0.01, n_obs_min=nom, want_host=False); eng.synchronize(); t3 = time.perf_counter()
        uni = eng.univar_nbrs(); order = fw.target_order(uni)
        t4 = time.perf_counter()
        res = eng.si_HITON_PC(order, max_k=max_k, alpha=0.01, n_obs_min=nom, want_tpc=False); t5 = time.perf_counter()
        nt = int(res.num_tests.sum())
        r = {"h2d_ms": (t1 - t0) * 1e3, "cor_ms": (t2 - t1) * 1e3, "pairwise_ms": (t3 - t2) * 1e3, "hiton_ms": (t5 - t4) * 1e3,
             "hiton_kernel_ms": eng.last_timing()["hiton_ms"], "cond_tests_ref": nt, "cond_tests_executed": res.tests_executed,
             "pairs": x.shape[0] * (x.shape[0] - 1) // 2, "nbr_entries": int(uni.offsets[-1]), "total_ms": (t5 - t0) * 1e3}
        if best is None or r["total_ms"] < best["total_ms"]:
            best = r
    out.update(best)
    out["cond_tests_per_s"] = out["cond_tests_ref"] / (out["hiton_ms"] * 1e-3) if out["cond_tests_ref"] else 0.0
    out["pairwise_tests_per_s"] = out["pairs"] / (out["pairwise_ms"] * 1e-3)
    print(json.dumps(out), flush=True)

which = sys.argv[1:] or ["C1", "C2", "C3", "C5s"]
if "C1" in which:
    run("C1 1000x500 fz max_k=0", synth.clique(1000, 500, B=24, seed=synth.BASE_SEED + 0), "fz", 0)
if "C2" in which:
    run("C2 10000x2000 fz max_k=3", synth.clique(10000, 2000, B=24, seed=synth.BASE_SEED + 1), "fz", 3)
if "C3" in which:
    run("C3 10000x2000 mi max_k=3", synth.binarize(synth.clique(10000, 2000, B=24, seed=synth.BASE_SEED + 2)), "mi", 3)
if "C3c" in which:
    run("C3 (chain) 10000x2000 mi max_k=3", synth.binarize(synth.chain(10000, 2000, B=32, seed=synth.BASE_SEED + 2)), "mi", 3)
if "C5s" in which:
    lat = synth.clique(4800, 10000, B=24, seed=synth.BASE_SEED + 4)
    run("C5-shaped (reduced p) 4800x10000 fz_nz max_k=3", synth.with_zeros(lat, zero_frac=0.4, seed=7), "fz_nz", 3, reps=2)
if "C5m" in which:
    x, _ = synth.hetero(9610, 10000, B=24, seed=synth.BASE_SEED + 4)
    run("C5 (reduced p) hetero+meta 9610x10000 fz_nz max_k=3", x, "fz_nz", 3, reps=2)
if "C5" in which:
    x, _ = synth.hetero(50010, 10000, B=24, seed=synth.BASE_SEED + 4)
    run("C5 hetero+meta 50010x10000 fz_nz max_k=3", x, "fz_nz", 3, reps=1)
if "PREP" in which:
    # normalisation throughput at the C4 shape: synthetic counts (Poisson around a log-normal depth), all six modes
    rng = np.random.default_rng(7)
    p_, n_ = 50000, 10000
    counts = np.empty((p_, n_), np.float32)
    for b in range(0, p_, 5000):
        lam = rng.gamma(0.3, 30.0, size=(5000, 1)).astype(np.float32) * rng.lognormal(0.0, 0.5, size=(1, n_)).astype(np.float32)
        counts[b:b + 5000] = rng.poisson(lam * (rng.random((5000, n_)) > 0.5)).astype(np.float32)
    eng = fw.Engine(0)
    rmask = np.zeros(n_, np.uint8); cmask = np.zeros(p_, np.uint8)
    import ctypes as C
    for mode in ("tss", "clr-nonzero", "clr-adapt", "pres-abs", "clr-nonzero-binned"):
        best = None
        for rep in range(2):
            n_out, p_out = C.c_int64(0), C.c_int64(0)
            t0 = time.perf_counter()
            eng._ck(eng.L.fw_normalize_f32(eng.h, counts.ctypes.data_as(C.c_void_p), n_, p_, n_, fw.NORM_MODES[mode], 3, C.byref(n_out), C.byref(p_out),
                                           rmask.ctypes.data_as(C.c_void_p), cmask.ctypes.data_as(C.c_void_p)))
            eng.synchronize(); dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        print(json.dumps({"config": "normalisation %s, %d x %d counts from pageable host memory" % (mode, p_, n_), "ms": best * 1e3,
                          "n_out": n_out.value, "p_out": p_out.value, "input_GBps": p_ * n_ * 4 / best / 1e9}), flush=True)
