"""Cycle stamps of cor_tc3_kernel per cluster (experiment build with -DFW_COR3_DEBUG, FW_LIB_PATH=build/exp/libfwgpu_dbg.so)."""
import os, sys, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fwload
fw = fwload.load(); synth = fwload.load_sub("synth")
p, n = int(sys.argv[1]), int(sys.argv[2])
x = synth.clique(p, n, B=16, seed=7)
eng = fw.Engine(0); eng.set_data_colmajor(x, "fz")
eng.cor(want_host=False); eng.synchronize()
ncl = 40000
buf = torch.zeros(ncl * 8, dtype=torch.int64, device="cuda")
assert eng.L.fw_debug_cor3_trace(C.c_void_p(buf.data_ptr())) == 0
eng.cor(want_host=False); eng.synchronize()
print("cor_ms", eng.last_timing()["cor_ms"])
d = buf.cpu().numpy().reshape(ncl, 8)
d = d[d[:, 0] != 0]
print("clusters traced", len(d))
ent, first, last_issue, t_s1, t_s2, mma_done, epi_done, exit_ = [d[:, i].astype(np.float64) for i in range(8)]
def s(name, v): print("%-34s mean %9.0f  p10 %9.0f  p50 %9.0f  p90 %9.0f" % (name, v.mean(), np.percentile(v, 10), np.percentile(v, 50), np.percentile(v, 90)))
s("entry -> first stage full", first - ent)
s("first full -> last MMA issued", last_issue - first)
s("last issue -> MMAs complete", mma_done - last_issue)
s("MMAs complete -> epilogue done", epi_done - mma_done)
lv = t_s2 > 0
s("  clamp + stage (warp 2)", (t_s1 - mma_done))
s("  mirror loop (live tiles)", (t_s2 - t_s1)[lv])
s("  direct loop (live tiles)", (epi_done - t_s2)[lv])
s("epilogue done -> cluster exit", exit_ - epi_done)
s("total entry -> exit", exit_ - ent)
