import sys; sys.path.insert(0, "/root/repo")
import numpy as np
import bmm_mcmc_b200 as B
from bmm_mcmc_b200.rcompat import RRng
from oracle import pyoracle as O
K, N, P = 72, 300, 6
rng = np.random.default_rng(9)
X = (rng.random((N, P)) < 0.5).astype(np.int32)
iz = RRng(3).sample_int(K, N)
r = O.gibbs_collapsed(X, iz, 6, K, burnin=0, seed=2, use_ref=False)
g = B.gibbs_collapsed(X, 6, K, burnin=0, initial_K=iz, replay={"u": r["u_rec"][None], "alpha": np.asfortranarray(r["alpha"])}, probes=("probs",))
mm = np.argwhere(g["z"] != r["z"])
j, i = mm[0]
print("first mismatch sweep", j, "obs", i, "gpu z", g["z"][j, i], "oracle z", r["z"][j, i])
pr_o, pr_g = r["probs"][j][i], g["probs"][j][i]
print("prob diff at that draw", np.abs(pr_o - pr_g).max())
u = r["u_rec"][j, i]
def walk(prob, dtype):
    p_tot = dtype(0)
    for v in prob: p_tot += dtype(v)
    slot = 0
    for k in range(K - 1):
        pk = prob[k]
        if pk != 0:
            pp = float(dtype(pk) / p_tot)
            if pp < 1:
                p_ = min(pp, 1 - pp); q = 1 - p_
                uu = u[slot]; slot += 1
                ix = 0 if uu < q else 1
                got = (1 - ix) if pp > 0.5 else ix
                if got:
                    return k + 1, uu, q, pp
            else:
                return k + 1, None, None, pp
        p_tot -= dtype(pk)
    return K, None, None, None
print("double walk     ", walk(pr_o, np.float64))
print("long double walk", walk(pr_o, np.longdouble))
