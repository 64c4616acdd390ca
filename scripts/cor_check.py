"""Bring-up check of the tensor-core cor_mat kernel against numpy fp64 (run on the GPU box)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fwload
fw = fwload.load(); synth = fwload.load_sub("synth")
for (p, n) in [(128, 64), (200, 346), (1000, 2000), (4096, 4096)]:
    x = synth.clique(p, n, B=16, seed=p)
    eng = fw.Engine(0)
    eng.set_data_colmajor(x, "fz")
    t0 = time.perf_counter(); got = eng.cor(); t1 = time.perf_counter()
    want = np.corrcoef(x.astype(np.float64))
    err = np.abs(got - want)
    print("p=%d n=%d max|err|=%.3e rms=%.3e  sym=%s diag1=%s time=%.1f ms (incl. D2H) dev=%.2f ms" % (p, n, err.max(), np.sqrt((err ** 2).mean()), (got == got.T).all(), (np.diag(got) == 1).all(), (t1 - t0) * 1e3, eng.last_timing()["cor_ms"]), flush=True)
    if err.max() > 1e-4:
        i, j = np.unravel_index(err.argmax(), err.shape)
        print("  worst at", i, j, got[i, j], want[i, j]); print(got[:4, :4]); print(want[:4, :4])
