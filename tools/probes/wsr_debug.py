"""Which rows of Q_final differ between a 148-CTA run and a capped-grid run of the single-pass relabelling (debug probe)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bmm_mcmc_b200 as B
rng = np.random.default_rng(99)
N, P, K = 20_000 + 77, 64, 32
th_true = rng.uniform(0.1, 0.9, (8, P))
X = (rng.random((N, P)) < th_true[rng.integers(0, 8, N)]).astype(np.int32)
cap = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for ns in (8, 9, 13):
    kw = dict(alpha=1.0, burnin=6, relabel=True, burnrelabel=3, seed=5, precision="fp32", probes=("Q_final",), grid_path=True)
    os.environ.pop("BMM_GRID_MAX_CTAS", None)
    ref = B.gibbs_stickbreaking(X, ns, K, **kw)
    os.environ["BMM_GRID_MAX_CTAS"] = str(cap)
    g = B.gibbs_stickbreaking(X, ns, K, **kw)
    bad = ~np.isclose(g["Q_final"], ref["Q_final"], rtol=3e-4, atol=0).all(axis=1)
    rows = np.nonzero(bad)[0]
    tiles = np.unique(rows // 128)
    print("ns", ns, "perm equal", np.array_equal(g["permutations"], ref["permutations"]), "bad rows", rows.size, "tiles", tiles.size,
          "tile % cap", np.unique(tiles % cap), "k = tile // cap", np.unique(tiles // cap)[:40])
    if rows.size:
        r = rows[0]
        print("  row", r, "got", g["Q_final"][r, :6], "want", ref["Q_final"][r, :6])
