"""Error structure of the tensor-core cor_mat (vs fp64) as a function of |r| and n."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fwload
fw = fwload.load(); synth = fwload.load_sub("synth")
for n in [1000, 4000, 10000]:
    p = 1024
    x = synth.clique(p, n, B=16, seed=n)
    rng = np.random.default_rng(1)
    for j, eps in enumerate([0.02, 0.1, 0.3, 1.0, 3.0]):
        x[16 * (j + 1) + 1] = x[16 * (j + 1)] + eps * rng.standard_normal(n).astype(np.float32)
    x[200] = -x[100] + 0.05 * rng.standard_normal(n).astype(np.float32)
    eng = fw.Engine(0); eng.set_data_colmajor(x, "fz")
    got = eng.cor().astype(np.float64)
    want = np.corrcoef(x.astype(np.float64))
    want32 = want.astype(np.float32).astype(np.float64)
    err = got - want
    iu = np.triu_indices(p, 1)
    e, r = err[iu], want[iu]
    print("n=%d: max|err|=%.2e rms=%.2e mean(err*sign(r))=%.2e  (fp32 rounding of exact alone: %.1e)" % (n, np.abs(e).max(), np.sqrt((e**2).mean()), (e*np.sign(r)).mean(), np.abs(want32-want).max()))
    for lo, hi in [(0, 0.05), (0.05, 0.3), (0.3, 0.7), (0.7, 0.95), (0.95, 1.01)]:
        m = (np.abs(r) >= lo) & (np.abs(r) < hi)
        if m.any():
            print("   |r| in [%.2f,%.2f): count %7d  max|err| %.2e  mean signed (toward zero<0) %.2e" % (lo, hi, m.sum(), np.abs(e[m]).max(), (e[m]*np.sign(r[m])).mean()))
