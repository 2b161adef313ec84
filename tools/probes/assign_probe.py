import ctypes as C, time, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from bmm_mcmc_b200 import _lib
L = _lib.lib()
dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
def run(name, cost):
    K = cost.shape[0]
    cf = np.asfortranarray(cost); perm = np.zeros(K, dtype=np.int32)
    L.bmm_assign_warp(K, dp(cf), ip(perm))
    t0 = time.perf_counter()
    for _ in range(5): L.bmm_assign_warp(K, dp(cf), ip(perm))
    dt = (time.perf_counter() - t0) / 5
    print("%-40s K=%d  %.2f ms  cost=%.6g perm_is_perm=%s" % (name, K, dt * 1e3, cost[perm, np.arange(K)].sum(), sorted(perm) == list(range(K))))
rng = np.random.default_rng(0)
K = 128
run("random uniform", rng.uniform(0, 1000, (K, K)))
c = np.tile(rng.uniform(1e7, 2e7, K), (K, 1)); c[np.arange(K), np.arange(K)] -= 1e6
run("column-constant + strong diagonal", c)
c = np.tile(rng.uniform(1e7, 2e7, K), (K, 1)); pm = rng.permutation(K); c[pm, np.arange(K)] -= 1e6
run("column-constant + permuted minimum", c)
run("all equal", np.full((K, K), 3.0e7))
c = np.tile(rng.uniform(1e7, 2e7, K), (K, 1)); c[:, :40] = 1.5e7
run("40 identical columns", c)
# one-hot world: N points, z true, p one-hot, Q from slightly different clustering
N = 200000; z = rng.integers(0, K, N); zq = z.copy(); flip = rng.random(N) < 0.05; zq[flip] = rng.integers(0, K, flip.sum())
Pm = np.full((N, K), 1e-6, dtype=np.float32); Pm[np.arange(N), z] = 1.0
Q = np.full((N, K), 1e-6, dtype=np.float32); Q[np.arange(N), zq] = 1.0
G = np.log(Q.astype(np.float64)).T @ Pm.astype(np.float64); s = (Pm * np.log(Pm)).sum(0)
run("one-hot batch-like cost", s[None, :] - G)
half = K // 2
Pm2 = Pm.copy(); Pm2[:, half:] = 1e-6   # half the clusters empty
G = np.log(Q.astype(np.float64)).T @ Pm2.astype(np.float64); s = (Pm2 * np.log(Pm2)).sum(0)
run("one-hot, half the sample columns empty", s[None, :] - G)
