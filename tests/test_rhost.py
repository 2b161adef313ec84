"""The Rcpp host of the R package (r-package/src/host.cpp), compiled unmodified against the Rcpp stand-in (oracle/shim/,
the one that also compiles the reference's own sources) and linked with libbmm_b200.so.  There is no R in the image; this
is how the host's argument forwarding, the names / shapes / storage modes of the list it returns
(/root/reference/src/full_gibbs.cpp:233-248, collapsed_gibbs.cpp:229-243) and its error propagation are exercised.
CPU: it builds, links and fails loudly without a device.  GPU: every element equals the ctypes path's for the same key."""
import ctypes as C
import os

import numpy as np
import pytest

import bmm_mcmc_b200 as B
from bmm_mcmc_b200 import _lib
from bmm_mcmc_b200.rcompat import RRng
from conftest import ROOT, gpu_available

def _build():
    import importlib.util
    spec = importlib.util.spec_from_file_location("rhost_build", os.path.join(ROOT, "tests", "rhost", "build.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    _lib.lib()      # libbmm_b200.so must exist
    L = C.CDLL(m.build())
    L.rhost_last_error.restype = C.c_char_p
    return L


def _call(L, sampler, X, K, ns, burnin, relabel, br, ip=None, th=None, iz=None, alpha=0.0, seed=7):
    N, P = X.shape
    S = ns - burnin
    Xf = np.asfortranarray(X, dtype=np.int32)
    out = dict(pi=np.full(S * K, np.nan), alpha=np.full(S, np.nan), permutations=np.full(S * K, -1, np.int32),
               z=np.zeros(S * N, np.int32), theta=np.full(K * P * S, np.nan), z_original=np.zeros(S * N, np.int32),
               theta_original=np.full(K * P * S, np.nan))
    names = C.create_string_buffer(256)
    dims = (C.c_int * 3)()
    p = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
    L.rhost_set_seed(C.c_uint(seed))
    rc = L.rhost_gibbs(sampler, p(Xf), N, P, p(ip), p(np.asfortranarray(th) if th is not None else None), p(iz), ns, K,
                       C.c_double(alpha), C.c_double(0.5), C.c_double(0.5), C.c_double(1.0), C.c_double(1.0), burnin, int(relabel), br,
                       names, 256, p(out["pi"]), p(out["alpha"]), p(out["permutations"]), p(out["z"]), p(out["theta"]),
                       p(out["z_original"]), p(out["theta_original"]), dims)
    return rc, names.value.decode().split(","), out, list(dims)


def _key(seed):
    """philox_seed() of the host: two uniforms of R's generator after set.seed(seed)"""
    u = RRng(seed).runif(2)
    return (int(np.floor(u[0] * 4294967296.0)) << 32) | int(np.floor(u[1] * 4294967296.0))


def test_rcpp_host_builds_against_the_stand_in():
    L = _build()
    for s in ("rhost_gibbs", "rhost_predictive", "rhost_lpsolve"):
        assert hasattr(L, s)


@pytest.mark.skipif(gpu_available(), reason="only meaningful on a box without a GPU")
def test_rcpp_host_turns_the_missing_device_into_an_r_error(datasets):
    """BMM_ERR_CUDA -> Rcpp::stop(bmm_last_error()), the R condition a user would see; no silent fallback."""
    L = _build()
    X = datasets["K2_N100_P5"]
    rng = RRng(3)
    ip = np.exp(rng.runif(2)); ip /= ip.sum()
    th = rng.runif(10).reshape(5, 2).T.copy()
    rc, _, _, _ = _call(L, 0, X, 2, 30, 5, False, 0, ip=ip, th=th)
    assert rc == -1 and b"CUDA" in L.rhost_last_error()


@pytest.mark.gpu
@pytest.mark.parametrize("sampler,relabel", [(0, True), (0, False), (1, True), (2, True), (3, False)])
def test_rcpp_host_equals_ctypes_path(datasets, sampler, relabel):
    if not gpu_available():
        pytest.skip("needs a GPU")
    L = _build()
    X = datasets["K3_N1000_P5"] if sampler != 3 else datasets["K2_N100_P5"]
    N, P = X.shape
    K, ns, burnin, br, seed = (3, 60, 20, 5, 11) if sampler != 3 else (12, 40, 10, 4, 5)
    rng = RRng(99)
    ip = np.exp(rng.runif(K)); ip /= ip.sum()
    th = rng.runif(K * P).reshape(P, K).T.copy()
    iz = RRng(4).sample_int(K, N).astype(np.int32)
    rc, names, out, dims = _call(L, sampler, X, K, ns, burnin, relabel, br, ip=ip, th=th, iz=iz, seed=seed)
    assert rc == 0, L.rhost_last_error()
    kw = dict(burnin=burnin, relabel=relabel, burnrelabel=br, seed=_key(seed))
    g = {0: lambda: B.gibbs_full(X, ns, K, initial_pi=ip, initial_theta=th, **kw),
         1: lambda: B.gibbs_stickbreaking(X, ns, K, initial_pi=ip, initial_theta=th, **kw),
         2: lambda: B.gibbs_collapsed(X, ns, K, initial_K=iz, **kw),
         3: lambda: B.gibbs_dp(X, ns, maxK=K, **kw)}[sampler]()
    S = ns - burnin
    want = (["pi"] if sampler <= 1 else []) + ["alpha", "permutations", "z", "theta"] + (["z_original", "theta_original"] if relabel else [])
    assert names == want                                   # the reference's list layout, in its order
    assert dims == [K, P, S]                               # theta is a K x P x S array
    assert np.array_equal(out["z"].reshape(N, S).T, g["z"])
    assert np.array_equal(out["theta"].reshape(S, P, K).transpose(2, 1, 0), g["theta"], equal_nan=True)
    assert np.array_equal(out["alpha"], np.asarray(g["alpha"]).reshape(-1))
    if sampler <= 1:
        assert np.array_equal(out["pi"].reshape(K, S).T, g["pi"])
    if relabel:
        assert np.array_equal(out["permutations"].reshape(K, S).T, g["permutations"])
        assert np.array_equal(out["z_original"].reshape(N, S).T, g["z_original"])
        assert np.array_equal(out["theta_original"].reshape(S, P, K).transpose(2, 1, 0), g["theta_original"], equal_nan=True)


@pytest.mark.gpu
def test_rcpp_host_predictive_and_lpsolve(datasets):
    if not gpu_available():
        pytest.skip("needs a GPU")
    L = _build()
    X = datasets["K3_N1000_P5"]
    fit = B.gibbs_full(X, 80, 3, burnin=20, seed=3)
    new = np.asfortranarray(X[:40], dtype=np.int32)
    K, P, S = fit["theta"].shape
    th = np.ascontiguousarray(fit["theta"].transpose(2, 1, 0)); pi = np.ascontiguousarray(fit["pi"].T)
    lp = np.zeros(40); mem = np.zeros(40 * K)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    assert L.rhost_predictive(p(new), 40, P, K, S, p(th), p(pi), p(lp), p(mem)) == 0, L.rhost_last_error()
    want = B.predictive(fit, X[:40])
    assert np.array_equal(lp, want["log_pred"]) and np.array_equal(mem.reshape(K, 40).T, want["membership"])
    cost = np.asfortranarray(np.random.default_rng(1).uniform(0, 9, (6, 6)))
    sol = np.zeros(36, np.int32)
    assert L.rhost_lpsolve(p(cost), 6, p(sol)) == 0
    sol = sol.reshape(6, 6).T                               # column-major K x K 0/1
    assert (sol.sum(0) == 1).all() and (sol.sum(1) == 1).all()
    from scipy.optimize import linear_sum_assignment
    r, c = linear_sum_assignment(cost)
    assert abs((cost * sol).sum() - cost[r, c].sum()) < 1e-9
