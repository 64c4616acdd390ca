"""One or two passes of the pipeline of a BASELINE configuration (used under ncu; not the bench).
usage: python scripts/probe_config.py C2|C3|C4|C5|C4bin [reps] [max_targets]     (C4bin: the C4 shape binarised - 63 MB of planes, not L2-resident per SM slice)"""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fwload
fw = fwload.load(); synth = fwload.load_sub("synth")
cfg = sys.argv[1]; reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1; max_t = int(sys.argv[3]) if len(sys.argv) > 3 else 0
S = synth.BASE_SEED
if cfg == "C2": x, kind = synth.clique(10000, 2000, B=24, seed=S + 1), "fz"
elif cfg == "C3": x, kind = synth.binarize(synth.clique(10000, 2000, B=24, seed=S + 2)), "mi"
elif cfg == "C4": x, kind = synth.clique(50000, 10000, B=24, seed=S + 3), "fz"
elif cfg == "C4bin": x, kind = synth.binarize(synth.clique(50000, 10000, B=24, seed=S + 3)), "mi"
elif cfg == "C5": x, kind = synth.hetero(50010, 10000, B=24, seed=S + 4)[0], "fz_nz"
elif cfg == "C5s": x, kind = synth.hetero(9610, 10000, B=24, seed=S + 4)[0], "fz_nz"
else: raise SystemExit("unknown config")
p, n = x.shape
nom = fw.auto_n_obs_min(kind, 3, 5, max_level=2) if kind == "mi" else 20
eng = fw.Engine(0)
for rep in range(reps):
    t0 = time.perf_counter(); eng.set_data_colmajor(x, kind); eng.synchronize(); t1 = time.perf_counter()
    if kind == "fz":
        eng.cor(want_host=False); eng.synchronize()
    t2 = time.perf_counter()
    eng.pw_univar_neighbors(alpha=0.01, n_obs_min=nom, want_host=False); eng.synchronize(); t3 = time.perf_counter()
    uni = eng.univar_nbrs(); order = fw.target_order(uni)
    if max_t: order = order[-max_t:]
    t4 = time.perf_counter()
    res = eng.si_HITON_PC(order, max_k=3, alpha=0.01, n_obs_min=nom, want_tpc=False); t5 = time.perf_counter()
    nt = int(res.num_tests.sum()); lt = eng.last_timing()
    print("%s rep %d: h2d %.1f | cor %.1f | pairwise %.1f (dev %.2f) | hiton %.1f (dev %.2f) ms: %d targets, %d ref tests -> %.3e tests/s (kernel %.3e) | exec_by_k %s"
          % (cfg, rep, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, lt["pairwise_ms"], (t5 - t4) * 1e3, lt["hiton_ms"], len(order), nt, nt / (t5 - t4),
             nt / (lt["hiton_ms"] * 1e-3), list(eng.hiton_exec_by_k())), flush=True)
