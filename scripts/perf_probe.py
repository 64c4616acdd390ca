"""Ad-hoc timing probe (not the bench): phases of the fz pipeline on a clique workload."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fwload
fw = fwload.load(); synth = fwload.load_sub("synth")
p, n, B = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
t0 = time.time(); x = synth.clique(p, n, B=B); print("gen %.2fs" % (time.time() - t0), flush=True)
eng = fw.Engine(0)
for rep in range(2):
    t0 = time.perf_counter(); eng.set_data_colmajor(x, "fz"); eng.synchronize(); t1 = time.perf_counter()
    eng.cor(want_host=False); eng.synchronize(); t2 = time.perf_counter()
    eng.pw_univar_neighbors(alpha=0.01, n_obs_min=20, want_host=False); eng.synchronize(); t3 = time.perf_counter()
    uni = eng.univar_nbrs(); t4 = time.perf_counter()
    order = fw.target_order(uni)
    res = eng.si_HITON_PC(order, max_k=3, alpha=0.01, n_obs_min=20, want_tpc=False); t5 = time.perf_counter()
    nt = int(res.num_tests.sum())
    print("rep %d: h2d %.1f ms | cor %.1f ms (%.1f TFLOP/s useful) | pairwise %.1f ms (%.2e pairs/s) | copy nbrs %.1f ms | hiton %.1f ms: %d ref tests, %d executed -> %.3e tests/s | stats %s | entries %d"
          % (rep, (t1 - t0) * 1e3, (t2 - t1) * 1e3, 2.0 * n * p * p / (t2 - t1) / 1e12, (t3 - t2) * 1e3, p * (p - 1) / 2 / (t3 - t2), (t4 - t3) * 1e3,
             (t5 - t4) * 1e3, nt, res.tests_executed, nt / (t5 - t4), eng.pairwise_stats(), uni.offsets[-1]), flush=True)
