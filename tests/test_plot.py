"""plot_gibbs (Python mirror of R/utils.R:114-209): argument list, defaults and the panel data."""
import inspect

import numpy as np

import bmm_mcmc_b200 as B


def _fake(S=12, N=40, K=3, P=4, seed=0):
    rng = np.random.default_rng(seed)
    z = rng.choice([1, 2], size=(S, N), p=[0.7, 0.3])
    z[5, :3] = 3                                   # cluster 3 appears once, below the threshold
    return {"z": z, "theta": rng.random((K, P, S)), "pi": rng.dirichlet(np.ones(K), S), "alpha": rng.random((S, 1))}


def test_signature_is_the_reference_one():
    sig = inspect.signature(B.plot_gibbs)
    names = list(sig.parameters)[:9]
    assert names == ["obj", "theta", "z", "pi", "heights", "cluster_threshold", "cluster_labels", "theta_labels", "theta_to_display"]
    d = {k: v.default for k, v in sig.parameters.items()}
    assert d["theta"] is True and d["z"] is True and d["pi"] is False and d["heights"] is None
    assert d["cluster_threshold"] == 0.1 and d["cluster_labels"] is None and d["theta_labels"] is None and d["theta_to_display"] is None


def test_panels():
    obj = _fake()
    p = B.plot_gibbs(obj, draw=False)
    assert set(p) == {"z", "theta"}                         # pi is off by default
    zp = p["z"]
    assert zp["sample"].min() == 2                          # the first sample is not drawn
    assert set(zp["cluster"]) == {1, 2}                     # cluster 3 never passes the 10 % threshold
    row = zp[(zp["sample"] == 4) & (zp["cluster"] == 1)].iloc[0]
    assert np.isclose(row["prop"], (obj["z"][3] == 1).mean())
    th = p["theta"]
    assert set(th["theta_var"]) == {1, 2, 3, 4}
    r = th[(th["sample"] == 7) & (th["cluster"] == 2) & (th["theta_var"] == 3)].iloc[0]
    assert r["value"] == obj["theta"][1, 2, 6]
    q = B.plot_gibbs(obj, pi=True, theta_to_display=["b", "d"], theta_labels=list("abcd"), cluster_labels=["x", "y", "w"],
                     draw=False)
    assert set(q) == {"pi", "z", "theta"} and set(q["theta"]["theta_var"]) == {"b", "d"} and set(q["z"]["cluster"]) == {"x", "y"}
    assert len(q["pi"]) == 12 * 3
    hist, edges = B.plot_alpha(obj, draw=False)
    assert hist.sum() == 12 and np.isclose(edges[1] - edges[0], 0.1)
