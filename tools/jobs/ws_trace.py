"""Scratch: phase stamps of the tensor sweep kernel at a given N (rows on one GPU)."""
import ctypes as C, sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
import bench
from bmm_mcmc_b200 import _lib, api
N = int(sys.argv[1])
X = bench.synth_rows(0, N, 64, 32); ip, th = bench.grid_init(32, 64)
plan = api.Plan(_lib.SAMPLER_STICKBREAKING, X, 12, 32, chains=1, seed=1, init_pi=ip, init_theta=th, precision="fp32", compact_z=True,
                grid_path=True, alpha=1.0, beta=0.5, gamma=0.5, a=1.0, b=1.0, burnin=1, relabel=False, burnrelabel=0)
L = _lib.lib()
for r in range(3):
    plan.run(); plan.sync()
    t = (C.c_uint64 * 8)(); L.bmm_debug_ws_trace(t)
    t = [int(x) for x in t]
    print("N=%d run %d: prologue %.1f us, first tile +%.1f, tiles %.1f, flush +%.1f, exit +%.1f, total %.1f us; kernel_ms %s" % (
        N, r, (t[1]-t[0])/1e3, (t[2]-t[1])/1e3, (t[3]-t[2])/1e3, (t[4]-t[3])/1e3, (t[5]-t[4])/1e3, (t[5]-t[0])/1e3, plan.kernel_ms()[0]/11))
plan.close()
