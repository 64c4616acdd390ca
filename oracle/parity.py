"""Sampled GPU-vs-oracle parity of si_HITON_PC at sizes where the oracle cannot run the whole table (TEST INFRASTRUCTURE:
imported by tests/ and by bench.py's parity / cpu_baseline legs only, never by the product).

For a sample of targets the oracle re-runs the reference's per-target loop (src/hiton.jl:283-400, "single" semantics) on
exactly the inputs the engine used: the engine's univariate neighbour lists, and - for the Fisher-z kind - the engine's own
Float32 cor_mat restricted to the variables involved (fw_cor_gather), so that the comparison is bit-exact on the statistic.
Only the sampled targets and their univariate neighbours are handed to the oracle (a few thousand variables)."""
import numpy as np

from . import fwo


def sampled_hiton_parity(eng, kind, table_pn, sample_pos, res, uni, max_k, alpha, n_obs_min, hps=5, n_threads=0, max_tests=10_000_000):
    """eng: the engine that produced `res` (HitonResult) from the neighbour lists `uni` (NbrCSR over all p variables);
    table_pn: the [p, n] host table (any object supporting fancy row indexing); sample_pos: positions in res.targets to check.
    Returns dict(targets, mismatches, cond_tests, max_stat_diff, max_p_rel, oracle_secs)."""
    import time
    sample_pos = np.asarray(sample_pos, np.int64)
    tg = np.asarray(res.targets, np.int64)[sample_pos]
    off = uni.offsets
    lists = [uni.nbr[off[T]:off[T + 1]] for T in tg]
    U = np.unique(np.concatenate([tg] + lists)) if len(tg) else np.zeros(0, np.int64)
    sub = np.ascontiguousarray(np.asarray(table_pn[U]))                    # [|U|, n]
    ora = fwo.Oracle(sub.T, kind)
    if kind == "fz":
        ora.set_cor(eng.cor_gather(U).astype(np.float64))                   # the engine's own Float32 cor_mat
    loc_off = np.zeros(len(tg) + 1, np.int64)
    loc_off[1:] = np.cumsum([len(l) for l in lists])
    cat = np.concatenate(lists) if loc_off[-1] else np.zeros(0, np.int64)
    loc_nbr = np.searchsorted(U, cat)
    sel = np.concatenate([np.arange(off[T], off[T + 1]) for T in tg]) if loc_off[-1] else np.zeros(0, np.int64)
    t0 = time.perf_counter()
    pcc, pn, ps, pp, nt = ora.hiton_pc_batch(np.searchsorted(U, tg), loc_off, loc_nbr, uni.stat[sel], uni.pval[sel], max_k=max_k, alpha=alpha,
                                             hps=hps, n_obs_min=n_obs_min, max_tests=max_tests, n_threads=n_threads)
    secs = time.perf_counter() - t0
    exact = kind in ("fz", "fz_nz")
    bad, max_ds, max_dp, first = 0, 0.0, 0.0, None
    for j, i in enumerate(sample_pos):
        gn, gs, gp = res.pc(int(i))
        a = loc_off[j]
        wn, ws, wp = U[pn[a:a + pcc[j]]], ps[a:a + pcc[j]], pp[a:a + pcc[j]]
        ok = len(gn) == len(wn) and (np.asarray(gn) == wn).all() and int(res.num_tests[int(i)]) == int(nt[j])
        if ok and len(gn):
            ds = np.abs(np.asarray(gs) - ws)
            dp = np.abs(np.asarray(gp) - wp) / np.maximum(np.abs(wp), 1e-300)
            max_ds = max(max_ds, float(ds.max())); max_dp = max(max_dp, float(dp[np.abs(wp) > 1e-290].max(initial=0.0)))
            ok = bool((ds == 0).all()) if exact else bool((ds <= 1e-12 * np.maximum(1.0, np.abs(ws))).all())
            ok = ok and bool((dp[np.abs(wp) > 1e-290] <= (1e-12 if exact else 1e-9)).all())
        if not ok:
            bad += 1
            if first is None:
                first = {"target": int(tg[j]), "gpu": [list(map(int, gn)), list(map(float, gs))], "oracle": [list(map(int, wn)), list(map(float, ws))],
                         "num_tests": [int(res.num_tests[int(i)]), int(nt[j])]}
    return {"targets": int(len(tg)), "mismatches": int(bad), "cond_tests": int(nt.sum()), "variables_in_oracle": int(len(U)),
            "max_stat_diff": max_ds, "max_p_rel": max_dp, "oracle_secs": secs, "first_mismatch": first}
