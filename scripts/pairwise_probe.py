import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fwload
fw = fwload.load(); synth = fwload.load_sub("synth")
p, n = int(sys.argv[1]), int(sys.argv[2])
x = synth.clique(p, n, B=24, seed=3)
eng = fw.Engine(0); eng.set_data_colmajor(x, "fz"); eng.cor(want_host=False)
for it in range(3):
    t0 = time.perf_counter(); eng.pw_univar_neighbors(alpha=0.01, n_obs_min=20, want_host=False); eng.synchronize(); t1 = time.perf_counter()
    print("pairwise wall %.2f ms, dev %.2f ms, stats %s" % ((t1 - t0) * 1e3, eng.last_timing()["pairwise_ms"], eng.pairwise_stats()), flush=True)
