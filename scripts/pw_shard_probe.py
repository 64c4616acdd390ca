"""Timing of the sharded pairwise stage (fw_pairwise_partial / merge) at the C5 shape on one GPU: every emulated rank in turn."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fwload
fw = fwload.load(); synth = fwload.load_sub("synth")
p, n = int(sys.argv[1]), int(sys.argv[2])
x = synth.hetero(p, n, B=24, seed=synth.BASE_SEED + 4)[0]
eng = fw.Engine(0); eng.set_data_colmajor(x, "fz_nz")
for rep in range(2):
    t0 = time.perf_counter(); eng.pw_univar_neighbors(alpha=0.01, n_obs_min=20, want_host=False); t1 = time.perf_counter()
    print("full: wall %.1f ms, device %.1f ms, stats %s" % ((t1 - t0) * 1e3, eng.last_timing()["pairwise_ms"], eng.pairwise_stats()), flush=True)
for world in (2, 8):
    recs = []
    for r in range(world):
        t0 = time.perf_counter(); rec = eng.pairwise_partial(r, world, alpha=0.01, n_obs_min=20); t1 = time.perf_counter()
        print("world %d rank %d: wall %.1f ms, device %.1f ms, raw %d, reliable %d" % (world, r, (t1 - t0) * 1e3, eng.last_timing()["pairwise_ms"], len(rec["x"]), rec["n_reliable"]), flush=True)
        recs.append(rec)
    t0 = time.perf_counter(); eng.pairwise_merge(recs, alpha=0.01, want_host=False); t1 = time.perf_counter()
    print("world %d merge: wall %.1f ms, stats %s" % (world, (t1 - t0) * 1e3, eng.pairwise_stats()), flush=True)
