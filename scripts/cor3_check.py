"""cor_tc3_kernel (cta_group::2) vs cor_tc2_kernel: same matrix? error vs fp64, timing.  Usage: cor3_check.py [worker MODE P N OUT]"""
import os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "worker":
    mode, p, n, out = sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
    import fwload
    fw = fwload.load(); synth = fwload.load_sub("synth")
    x = synth.clique(p, n, B=16, seed=7)
    eng = fw.Engine(0); eng.set_data_colmajor(x, "fz")
    ts = []
    for it in range(4):
        eng.cor(want_host=False); eng.synchronize(); ts.append(eng.last_timing()["cor_ms"])
    print("mode %s p=%d n=%d cor_ms %s -> %.1f TFLOP/s useful" % (mode, p, n, ["%.2f" % v for v in ts], 2.0 * n * p * p / min(ts[1:]) / 1e9), flush=True)
    if p <= 8192:
        c = eng.cor()
        np.save(out, c)
        if p <= 2048:
            want = np.corrcoef(x.astype(np.float64))
            print("   max |err| vs fp64 %.2e, symmetric %s, unit diag %s" % (np.abs(c - want).max(), (c == c.T).all(), (np.diag(c) == 1).all()), flush=True)
        eng2 = fw.Engine(0)
        c2 = eng2.upload_and_cor(x, want_host=True)
        print("   upload-overlapped identical: %s" % (c2 == c).all(), flush=True)
    sys.exit(0)
shapes = [(50, 346), (333, 1000), (1024, 2000), (4096, 4096), (8192, 2048)] + ([(50000, 10000)] if "--big" in sys.argv else [])
for p, n in shapes:
    outs = {}
    for mode in ("1", "2"):
        out = "/tmp/cor3_%s.npy" % mode
        if os.path.exists(out): os.remove(out)
        env = dict(os.environ, FWGPU_COR_CLUSTER=mode)
        r = subprocess.run([sys.executable, __file__, "worker", mode, str(p), str(n), out], env=env, timeout=240)
        if r.returncode != 0:
            print("mode %s FAILED rc=%d" % (mode, r.returncode), flush=True)
        outs[mode] = np.load(out) if os.path.exists(out) else None
    if outs["1"] is not None and outs["2"] is not None:
        d = np.abs(outs["1"].astype(np.float64) - outs["2"].astype(np.float64))
        print("   tc2 vs tc3: identical %s, max |diff| %.2e, nan mismatch %d" % ((outs["1"] == outs["2"]).all(), np.nanmax(d), int((np.isnan(outs["1"]) != np.isnan(outs["2"])).sum())), flush=True)
