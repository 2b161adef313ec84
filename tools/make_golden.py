#!/usr/bin/env python3
"""Golden outputs of the CPU oracle on the bundled datasets (tests/golden/oracle_golden.npz).

The reference cannot be run here (no R), so these are regression vectors of the oracle
restatement (seeded with its R-compatible Mersenne-Twister), not reference outputs.  They keep the
restatement from drifting and give the GPU parity tests fixed targets.
    python tools/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bmm_mcmc_b200.rcompat import RRng  # noqa: E402


def _init_full(K, P, seed):
    rng = RRng(seed)
    ip = np.exp(rng.runif(K)); ip /= ip.sum()
    return ip, rng.runif(K * P).reshape(P, K).T


def cases(O, data):
    use_ref = O.has_ref()

    def full():
        X = data["K3_N1000_P5"]
        ip, th = _init_full(3, 5, 5)
        r = O.gibbs_full(X, ip, th, 30, 3, burnin=10, relabel=True, burnrelabel=4, seed=11, use_ref=use_ref)
        return {"z": r["z"], "z_rel": r["z_rel"], "perm": r["permutations"], "pi": r["pi"], "theta": r["theta"],
                "alpha": r["alpha"], "probs_last": r["probs"][-1], "loglik_last": r["loglik"][-1]}

    def collapsed():
        X = data["K2_N100_P5"]
        iz = RRng(3).sample_int(2, 100)
        r = O.gibbs_collapsed(X, iz, 40, 2, burnin=12, relabel=True, burnrelabel=5, seed=21, use_ref=use_ref)
        return {"z": r["z"], "z_rel": r["z_rel"], "perm": r["permutations"], "alpha": r["alpha"],
                "theta": r["theta"], "probs_last": r["probs"][-1], "Q": r["Q_final"]}

    def dp():
        X = data["K2_N1000_P5"]
        r = O.gibbs_dp(X, 25, maxK=64, seed=8, use_ref=use_ref)
        return {"z": r["z"], "alpha": r["alpha"], "Kactive": r["Kactive"], "probs_last": r["probs"][-1]}

    def stick():
        X = data["K3_N1000_P5"]
        ip, th = _init_full(8, 5, 9)
        r = O.gibbs_stickbreaking(X, ip, th, 25, 8, burnin=5, seed=4, use_ref=use_ref)
        return {"z": r["z"], "pi": r["pi"], "alpha": r["alpha"]}

    return {"full": full, "collapsed": collapsed, "dp": dp, "stick": stick}


def main():
    from oracle import pyoracle as O
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import load_golden_matrix
    data = {n: load_golden_matrix(n) for n in ("K2_N100_P5", "K2_N1000_P5", "K3_N1000_P5")}
    out = {}
    for name, fn in cases(O, data).items():
        for k, v in fn().items():
            out["%s__%s" % (name, k)] = v
    path = os.path.join(ROOT, "tests", "golden", "oracle_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
