"""cor_mat GEMM at the C4 shape with and without the epilogue collection of pairwise candidates; upload-overlapped variant."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fwload, torch
fw = fwload.load(); synth = fwload.load_sub("synth")
p, n = int(os.environ.get("P", 50000)), int(os.environ.get("N", 10000))
host = torch.empty((p, n), dtype=torch.float32, pin_memory=True)
host.numpy()[:] = synth.clique(p, n, B=24, seed=synth.BASE_SEED + 3)
eng = fw.Engine(0)
eng.set_data_ptr(host.data_ptr(), n, p)
for arm in (0.0, 0.01, 0.0, 0.01):
    eng.pairwise_prefetch(arm, 20)
    t = []
    for _ in range(4):
        eng.cor(want_host=False); eng.synchronize(); t.append(eng.last_timing()["cor_ms"])
    print("resident, prefetch alpha=%g: cor_ms %s" % (arm, ["%.2f" % x for x in t]), flush=True)
for arm in (0.0, 0.01, 0.0, 0.01):
    eng.pairwise_prefetch(arm, 20)
    t = []
    for _ in range(4):
        t0 = time.perf_counter(); eng.upload_and_cor(host.data_ptr(), n=n, p=p); eng.synchronize(); t.append((time.perf_counter() - t0) * 1e3)
    print("upload-overlapped, prefetch alpha=%g: wall ms %s" % (arm, ["%.2f" % x for x in t]), flush=True)
eng.pairwise_prefetch(0.01, 20)
eng.cor(want_host=False)
for _ in range(3):
    t0 = time.perf_counter(); eng.pw_univar_neighbors(alpha=0.01, n_obs_min=20, want_host=False); eng.synchronize()
    print("pairwise from collected list: wall %.2f ms, device %.2f" % ((time.perf_counter() - t0) * 1e3, eng.last_timing()["pairwise_ms"]), eng.pairwise_stats(), flush=True)
eng.pairwise_prefetch(0.0, 0)
eng.cor(want_host=False)
for _ in range(3):
    t0 = time.perf_counter(); eng.pw_univar_neighbors(alpha=0.01, n_obs_min=20, want_host=False); eng.synchronize()
    print("pairwise with scan: wall %.2f ms, device %.2f" % ((time.perf_counter() - t0) * 1e3, eng.last_timing()["pairwise_ms"]), flush=True)
