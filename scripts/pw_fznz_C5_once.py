import sys, os, time
import numpy as np
sys.path.insert(0, "/root/repo")
import fwload
fw = fwload.load(); synth = fwload.load_sub("synth")
x = synth.hetero(50010, 10000, B=24, seed=synth.BASE_SEED + 4)[0]
eng = fw.Engine(0); eng.set_data_colmajor(x, "fz_nz")
eng.pw_univar_neighbors(alpha=0.01, n_obs_min=20, want_host=False)
print(eng.last_timing()["pairwise_ms"], eng.pairwise_stats())
