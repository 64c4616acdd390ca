"""Warm timing of fw_cor_matrix (device events) at several sizes."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fwload
fw = fwload.load(); synth = fwload.load_sub("synth")
for (p, n) in [(4096, 4096), (8192, 2048), (16384, 10000)]:
    x = np.random.default_rng(0).standard_normal((p, n), dtype=np.float32)
    eng = fw.Engine(0); eng.set_data_colmajor(x, "fz")
    ts = []
    for it in range(4):
        eng.cor(want_host=False); eng.synchronize(); ts.append(eng.last_timing()["cor_ms"])
    t = min(ts[1:])
    print("p=%d n=%d cor_ms %s -> %.1f TFLOP/s useful (2np^2), tiles=%d" % (p, n, ["%.2f" % v for v in ts], 2.0 * n * p * p / t / 1e9, (p // 128) * (p // 128 + 1) // 2), flush=True)
