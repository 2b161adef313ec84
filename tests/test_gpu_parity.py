"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle.

Bars (BASELINE.json north_star): log-likelihood and conditional-probability matrices within 1e-6
relative (fp64); in replay mode (the kernels consume the oracle's recorded uniforms and parameter
draws) allocations, counts and permutations bit-exact.
"""
import numpy as np
import pytest

import bmm_mcmc_b200 as B
from bmm_mcmc_b200 import _lib
from bmm_mcmc_b200.rcompat import RRng
from conftest import gpu_available

pytestmark = pytest.mark.gpu

RTOL = 1e-6


def _need_gpu():
    # On the GPU box the library must be the thing that runs: fail loudly, never skip silently.
    assert gpu_available(), "CUDA library/device unavailable: " + repr(_lib.LIB_PATH)


def _init_full(K, P, seed):
    rng = RRng(seed)
    ip = np.exp(rng.runif(K))
    ip /= ip.sum()
    th = rng.runif(K * P).reshape(P, K).T  # K x P
    return ip, th


def _replay_of(r, keys=("pi", "theta", "alpha")):
    rp = {"u": r["u_rec"][None]}
    for k in keys:
        if k in r:
            rp[k] = np.asfortranarray(r[k])
    return rp


def _close(a, b, rtol=RTOL, atol=0.0):
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol)


def _counts(z, X, K):
    """c_k and V_kd from an S x N allocation history."""
    S = z.shape[0]
    ck = np.stack([(z == k + 1).sum(1) for k in range(K)], 1)
    V = np.stack([(z == k + 1).astype(np.int64) @ X for k in range(K)], 1)
    return ck, V


@pytest.mark.parametrize("name,K,relabel", [("K3_N1000_P5", 3, True), ("K2_N100_P5", 2, False), ("K2_N1000_P5", 2, True)])
def test_full_replay(oracle, datasets, name, K, relabel):
    _need_gpu()
    X = datasets[name]
    N, P = X.shape
    ns, burnin, br = 60, 20, 8
    ip, th = _init_full(K, P, 5)
    r = oracle.gibbs_full(X, ip, th, ns, K, burnin=burnin, relabel=relabel, burnrelabel=br, seed=11)
    g = B.gibbs_full(X, ns, K, burnin=burnin, relabel=relabel, burnrelabel=br, initial_pi=ip, initial_theta=th,
                     replay=_replay_of(r), probes=("probs", "loglik", "Q_final"))
    t = r.tail()
    # deterministic matrices
    _close(g["loglik"][1:], r["loglik"][1:])
    _close(g["probs"][1:], r["probs"][1:], atol=1e-300)
    # allocations, counts, permutations: bit-exact
    zo = "z_original" if relabel else "z"
    assert np.array_equal(g[zo], t[zo])
    ck_g, V_g = _counts(g[zo], X, K)
    ck_o, V_o = _counts(t[zo], X, K)
    assert np.array_equal(ck_g, ck_o) and np.array_equal(V_g, V_o)
    if relabel:
        assert np.array_equal(g["permutations"], t["permutations"])
        assert np.array_equal(g["z"], t["z"])
        _close(g["theta"], t["theta"], rtol=0)
        _close(g["Q_final"], r["Q_final"], rtol=1e-9)
    _close(g["theta_original" if relabel else "theta"], t["theta_original" if relabel else "theta"], rtol=0)
    _close(g["pi"], t["pi"], rtol=0)
    _close(g["alpha"], t["alpha"], rtol=0)


def test_stickbreaking_replay(oracle, datasets):
    _need_gpu()
    X = datasets["K3_N1000_P5"]
    N, P = X.shape
    K, ns, burnin, br = 8, 50, 20, 6
    ip, th = _init_full(K, P, 9)
    r = oracle.gibbs_stickbreaking(X, ip, th, ns, K, burnin=burnin, relabel=True, burnrelabel=br, seed=4)
    g = B.gibbs_stickbreaking(X, ns, K, burnin=burnin, relabel=True, burnrelabel=br, initial_pi=ip,
                              initial_theta=th, replay=_replay_of(r), probes=("probs",))
    t = r.tail()
    _close(g["probs"][1:], r["probs"][1:], atol=1e-300)
    assert np.array_equal(g["z_original"], t["z_original"])
    # unused sticks tie in the assignment: compare the labels that are occupied
    occ = np.unique(t["z_original"]) - 1
    assert np.array_equal(g["permutations"][:, occ], t["permutations"][:, occ])


def test_stickbreaking_burnrelabel_above_burnin_replay(oracle, datasets):
    """The reference's gibbs_stickbreaking wrapper does not clamp burnrelabel (R/utils.R:95-107): the C++ runs with
    probs_out slices that were never written (zero -> 1e-6 in my_stephens_batch).  Same here, chain and grid path."""
    _need_gpu()
    X = datasets["K2_N100_P5"]
    K, ns, burnin, br = 4, 60, 6, 50
    ip, th = _init_full(K, X.shape[1], 5)
    r = oracle.gibbs_stickbreaking(X, ip, th, ns, K, burnin=burnin, relabel=True, burnrelabel=br, seed=3)
    t = r.tail()
    occ = np.unique(t["z_original"]) - 1
    for grid in (False, True):
        g = B.gibbs_stickbreaking(X, ns, K, burnin=burnin, relabel=True, burnrelabel=br, initial_pi=ip, initial_theta=th,
                                  replay=_replay_of(r), grid_path=grid)
        assert np.array_equal(g["z_original"], t["z_original"]), grid
        assert np.array_equal(g["permutations"][:, occ], t["permutations"][:, occ]), grid


@pytest.mark.parametrize("sampler", ["full", "stickbreaking", "collapsed", "dp", "grid"])
def test_plan_rerun_is_identical(datasets, sampler):
    """bmm_plan_run restores the chain state bmm_plan_create built, so every run of a plan is the same chain
    (the DP sampler used to seat all observations again on top of the previous run's clusters)."""
    _need_gpu()
    from bmm_mcmc_b200 import api
    X = datasets["K2_N100_P5"]
    N, P = X.shape
    ns, burnin, br, C_ = 40, 12, 5, 3
    kw = dict(alpha=0.0, beta=0.5, gamma=0.5, a=1.0, b=1.0, burnin=burnin, relabel=True, burnrelabel=br, seed=11)
    if sampler in ("full", "stickbreaking", "grid"):
        K = 4
        C_ = 1 if sampler == "grid" else C_
        rng = RRng(2)
        ip = np.exp(rng.runif(C_ * K)).reshape(C_, K)
        ip /= ip.sum(1, keepdims=True)
        th = rng.runif(C_ * K * P).reshape(C_, P, K)
        sid = _lib.SAMPLER_STICKBREAKING if sampler == "stickbreaking" else _lib.SAMPLER_FULL
        plan = api.Plan(sid, X, ns, K, chains=C_, init_pi=np.ascontiguousarray(ip), init_theta=np.ascontiguousarray(th),
                        grid_path=sampler == "grid", **kw)
    elif sampler == "collapsed":
        iz = np.stack([RRng(5 + c).sample_int(3, N) for c in range(C_)]).astype(np.int32)
        plan = api.Plan(_lib.SAMPLER_COLLAPSED, X, ns, 3, chains=C_, init_z=np.ascontiguousarray(iz), **kw)
    else:
        plan = api.Plan(_lib.SAMPLER_DP, X, ns, 20, chains=C_, **kw)
    plan.run()
    first = {k: v.copy() for k, v in plan.fetch().items()}
    for _ in range(2):
        plan.run()
    again = plan.fetch()
    plan.close()
    assert set(first) == set(again)
    for k in first:
        assert np.array_equal(first[k], again[k], equal_nan=True), (sampler, k)


@pytest.mark.parametrize("kernel", ["fast", "generic"])
@pytest.mark.parametrize("name,K,relabel,alpha", [("K2_N100_P5", 2, False, 0.0), ("K3_N1000_P5", 3, True, 0.0),
                                                  ("K2_N1000_P5", 4, True, 1.5)])
def test_collapsed_replay(oracle, datasets, monkeypatch, name, K, relabel, alpha, kernel):
    """Both collapsed kernels: the register-resident low-latency one (few chains) and the generic one."""
    _need_gpu()
    monkeypatch.setenv("BMM_COLLAPSED_KERNEL", kernel)
    X = datasets[name]
    N, P = X.shape
    ns, burnin, br = 40, 12, 5
    iz = RRng(3).sample_int(K, N)
    r = oracle.gibbs_collapsed(X, iz, ns, K, alpha=alpha, burnin=burnin, relabel=relabel, burnrelabel=br, seed=21)
    g = B.gibbs_collapsed(X, ns, K, alpha=alpha if alpha else None, burnin=burnin, relabel=relabel, burnrelabel=br,
                          initial_K=iz, replay=_replay_of(r, keys=("alpha",)), probes=("probs", "Q_final"))
    t = r.tail()
    _close(g["probs"][1:], r["probs"][1:], atol=1e-300)
    zo = "z_original" if relabel else "z"
    assert np.array_equal(g[zo], t[zo])
    tho = "theta_original" if relabel else "theta"
    np.testing.assert_array_equal(g[tho][:, :, 1:] if burnin == 0 else g[tho], t[tho])  # S_kd / N_k exactly; NaN == NaN
    _close(g["alpha"], t["alpha"], rtol=0)
    if relabel:
        assert np.array_equal(g["permutations"], t["permutations"])
        assert np.array_equal(g["z"], t["z"])
        _close(g["Q_final"], r["Q_final"], rtol=1e-9)


@pytest.mark.parametrize("name,K", [("K2_N100_P5", 2), ("K3_N1000_P5", 3), ("K2_N1000_P5", 7), ("K3_N1000_P5", 20)])
def test_collapsed_product_form_equals_log_form(datasets, monkeypatch, name, K):
    """Philox mode runs the conditional of collapsed_gibbs.cpp:105-130 in product form (no exp, no log tables).
    Same seed, same uniforms: the probabilities equal the log-form kernel's to 1e-12 relative in double (the two
    differ only by rounding), to 1e-6 in single precision, and the chains are the same chains."""
    _need_gpu()
    X = datasets[name]
    kw = dict(burnin=4, relabel=True, burnrelabel=2, chains=3, seed=17, probes=("probs",))
    monkeypatch.setenv("BMM_COLLAPSED_KERNEL", "log")
    ref = B.gibbs_collapsed(X, 12, K, **kw)
    monkeypatch.delenv("BMM_COLLAPSED_KERNEL")
    g = B.gibbs_collapsed(X, 12, K, **kw)
    assert np.array_equal(g["z_original"], ref["z_original"]) and np.array_equal(g["permutations"], ref["permutations"])
    _close(g["probs"][:, 1:], ref["probs"][:, 1:], rtol=1e-12, atol=1e-300)
    _close(g["theta"], ref["theta"], rtol=0)
    f = B.gibbs_collapsed(X, 12, K, precision="fp32", **kw)
    # single precision: the first sweep starts from the same state, so its probabilities are comparable entry by entry
    _close(f["probs"][:, 1], ref["probs"][:, 1], rtol=2e-6, atol=1e-30)
    assert (f["z_original"][:, 0] != ref["z_original"][:, 0]).mean() < 1e-3


@pytest.mark.parametrize("name,maxK", [("K2_N100_P5", 30), ("K2_N1000_P5", 64)])
def test_dp_product_form_equals_log_form(datasets, monkeypatch, name, maxK):
    """The DP sampler's Philox mode evaluates the conditional of collapsed_gibbs_dp.cpp:140-186 in product form;
    same seed => same chains as the log / exp-normalise form, probabilities equal to rounding."""
    _need_gpu()
    X = datasets[name]
    kw = dict(burnin=4, relabel=False, maxK=maxK, chains=3, seed=23, probes=("probs",))
    monkeypatch.setenv("BMM_DP_LOGFORM", "1")
    ref = B.gibbs_dp(X, 10, **kw)
    monkeypatch.delenv("BMM_DP_LOGFORM")
    g = B.gibbs_dp(X, 10, **kw)
    assert np.array_equal(g["z"], ref["z"])
    _close(g["probs"][:, 1:], ref["probs"][:, 1:], rtol=1e-11, atol=1e-300)
    _close(g["alpha"], ref["alpha"], rtol=0)


@pytest.mark.parametrize("name,maxK,relabel", [("K2_N1000_P5", 64, False), ("K2_N100_P5", 30, True), ("K2_N100_P5", 4, False)])
def test_dp_replay(oracle, datasets, name, maxK, relabel):
    _need_gpu()
    X = datasets[name]
    N, P = X.shape
    ns, burnin, br = 40, 12, 5
    try:
        r = oracle.gibbs_dp(X, ns, alpha=0.0, burnin=burnin, relabel=relabel, burnrelabel=br, maxK=maxK, seed=8)
    except RuntimeError as e:
        # truncation drove the reference's state into undefined behaviour (quirk 9): the GPU must flag it too
        with pytest.raises(_lib.BmmError):
            B.gibbs_dp(X, ns, burnin=burnin, relabel=relabel, burnrelabel=br, maxK=maxK, seed=8)
        return
    g = B.gibbs_dp(X, ns, burnin=burnin, relabel=relabel, burnrelabel=br, maxK=maxK,
                   replay=_replay_of(r, keys=("alpha",)), probes=("probs",))
    t = r.tail()
    _close(g["probs"][1:], r["probs"][1:], atol=1e-300)
    zo = "z_original" if relabel else "z"
    assert np.array_equal(g[zo], t[zo])
    np.testing.assert_array_equal(g["theta_original" if relabel else "theta"], t["theta_original" if relabel else "theta"])
    if relabel:
        # Most of the maxK labels are unused, so the K x K assignment has exactly tied optima
        # (SURVEY 8c): lp_solve and the GPU solver may pick different ones.  Check what is
        # well-defined: each row is a permutation and z is z_original mapped through it.
        S = g["permutations"].shape[0]
        assert np.array_equal(np.sort(g["permutations"], 1), np.tile(np.arange(maxK), (S, 1)))
        assert np.array_equal(g["z"], np.take_along_axis(g["permutations"], g["z_original"] - 1, 1) + 1)


def test_condprob_probe(oracle, datasets):
    """One z-sweep's log-likelihood / conditional-probability matrices at a fixed state."""
    _need_gpu()
    import ctypes as C
    X = np.asfortranarray(datasets["K3_N1000_P5"])
    N, P = X.shape
    K = 3
    ip, th = _init_full(K, P, 2)
    r = oracle.gibbs_full(X, ip, th, 2, K, burnin=0, seed=1)
    ll = np.zeros((N, K), order="F")
    pr = np.zeros((N, K), order="F")
    thf = np.asfortranarray(th)
    L = _lib.lib()
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    rc = L.bmm_full_condprob(X.ctypes.data_as(C.POINTER(C.c_int32)), N, P, K, dp(thf), dp(ip), 0, 0, dp(ll), dp(pr))
    _lib.check(rc)
    _close(ll, r["loglik"][1])
    _close(pr, r["probs"][1])


def test_assign_matches_lpsolve(oracle):
    _need_gpu()
    import ctypes as C
    L = _lib.lib()
    rng = np.random.default_rng(0)
    for K in (2, 3, 4, 5, 6, 8, 16, 32):
        batch = 20
        cost = rng.uniform(0, 1000, (batch, K, K))
        cf = np.ascontiguousarray(np.stack([np.asfortranarray(c).ravel(order="F") for c in cost]))
        sol = np.zeros((batch, K * K), dtype=np.int32)
        _lib.check(L.bmm_assign(K, batch, cf.ctypes.data_as(C.POINTER(C.c_double)), sol.ctypes.data_as(C.POINTER(C.c_int32))))
        for b in range(batch):
            s_g = sol[b].reshape(K, K, order="F")
            s_o = oracle.assign(cost[b], use_ref=oracle.has_ref())
            assert (s_g.sum(0) == 1).all() and (s_g.sum(1) == 1).all()
            assert np.isclose((cost[b] * s_g).sum(), (cost[b] * s_o).sum(), rtol=1e-12)
            assert np.array_equal(s_g, s_o)  # random real costs: the optimum is unique


def test_stephens_helpers(oracle):
    _need_gpu()
    import ctypes as C
    L = _lib.lib()
    rng = np.random.default_rng(1)
    N, K, M = 200, 3, 7
    p = rng.dirichlet(np.ones(K) * 0.5, size=(M, N)).transpose(1, 2, 0)  # N x K x M
    p[3, 1, 2] = 0.0
    pf = np.asfortranarray(p)
    q_o, perm_o = oracle.stephens_batch(pf, use_ref=oracle.has_ref())
    q_g = np.zeros((N, K), order="F")
    perm_g = np.zeros((M, K), dtype=np.int32, order="F")
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
    _lib.check(L.bmm_stephens_batch(N, K, M, dp(pf), dp(q_g), ip(perm_g)))
    assert np.array_equal(perm_g, perm_o)
    _close(q_g, q_o, rtol=1e-12)
    ps = np.asfortranarray(rng.dirichlet(np.ones(K), size=N))
    perm2_o, qn_o, cost_o = oracle.stephens_online(q_o, ps, 17, use_ref=oracle.has_ref())
    perm2_g = np.zeros(K, dtype=np.int32)
    qn_g = np.zeros((N, K), order="F")
    cost_g = np.zeros((K, K), order="F")
    _lib.check(L.bmm_stephens_online(N, K, dp(np.asfortranarray(q_o)), dp(ps), 17, ip(perm2_g), dp(qn_g), dp(cost_g)))
    _close(cost_g, cost_o, rtol=1e-9)
    assert np.array_equal(perm2_g, perm2_o)
    _close(qn_g, qn_o, rtol=1e-12)


def test_stephens_fixed_mode_helpers(oracle):
    """BMM_FLAG_STEPHENS_FIXED through the helper entry points against the oracle's fixed mode (inverse permutation,
    log p online cost, running-mean Q): same permutations, Q and cost to rounding."""
    _need_gpu()
    import ctypes as C
    L = _lib.lib()
    rng = np.random.default_rng(8)
    N, K, M = 150, 4, 6
    base = rng.dirichlet(np.ones(K) * 0.3, N)
    sig = [rng.permutation(K) for _ in range(M)]
    p = np.stack([0.9 * base[:, sg] + 0.1 * rng.dirichlet(np.ones(K), N) for sg in sig], axis=2)
    p[5, 2, 1] = 0.0
    pf = np.asfortranarray(p)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
    try:
        oracle.set_stephens_fixed(True)
        q_o, perm_o = oracle.stephens_batch(pf, use_ref=oracle.has_ref())
        ps = np.asfortranarray(base[:, sig[0]])
        ps[7, 1] = 0.0
        perm2_o, qn_o, cost_o = oracle.stephens_online(q_o, ps, 17, use_ref=oracle.has_ref())
    finally:
        oracle.set_stephens_fixed(False)
    q_g = np.zeros((N, K), order="F")
    perm_g = np.zeros((M, K), dtype=np.int32, order="F")
    _lib.check(L.bmm_stephens_batch_ex(N, K, M, dp(pf), dp(q_g), ip(perm_g), _lib.FLAG_STEPHENS_FIXED))
    assert np.array_equal(perm_g, perm_o)
    _close(q_g, q_o, rtol=1e-12)
    perm2_g = np.zeros(K, dtype=np.int32)
    qn_g = np.zeros((N, K), order="F")
    cost_g = np.zeros((K, K), order="F")
    _lib.check(L.bmm_stephens_online_ex(N, K, dp(np.asfortranarray(q_o)), dp(ps), 17, ip(perm2_g), dp(qn_g), dp(cost_g),
                                        _lib.FLAG_STEPHENS_FIXED))
    _close(cost_g, cost_o, rtol=1e-9)
    assert np.array_equal(perm2_g, perm2_o)
    _close(qn_g, qn_o, rtol=1e-12)


@pytest.mark.parametrize("grid", [False, True])
def test_full_replay_stephens_fixed(oracle, datasets, grid):
    """gibbs_full with relabelling in the correctness-fixed Stephens mode, replayed against the oracle in the same
    mode, on the chain-per-block path (fp64) and the streaming grid path (float P / Q)."""
    _need_gpu()
    X = datasets["K3_N1000_P5"]
    N, P = X.shape
    K, ns, burnin, br = 3, 50, 20, 6
    ip, th = _init_full(K, P, 5)
    try:
        oracle.set_stephens_fixed(True)
        r = oracle.gibbs_full(X, ip, th, ns, K, burnin=burnin, relabel=True, burnrelabel=br, seed=11)
    finally:
        oracle.set_stephens_fixed(False)
    r0 = oracle.gibbs_full(X, ip, th, ns, K, burnin=burnin, relabel=True, burnrelabel=br, seed=11)
    g = B.gibbs_full(X, ns, K, burnin=burnin, relabel=True, burnrelabel=br, initial_pi=ip, initial_theta=th,
                     replay=_replay_of(r), probes=("Q_final",), grid_path=grid, stephens_fixed=True)
    t = r.tail()
    assert np.array_equal(g["z_original"], t["z_original"])
    assert np.array_equal(g["permutations"], t["permutations"])
    assert np.array_equal(g["z"], t["z"])
    _close(g["theta"], t["theta"], rtol=0)
    _close(g["Q_final"], r["Q_final"], rtol=2e-4 if grid else 1e-9)
    assert not np.allclose(r["Q_final"], r0["Q_final"])      # the mode does change the reference Q (running mean)


def test_chain_split_invariance(datasets):
    """Chains are keyed by their global index: running chains 2..3 alone equals rows 2..3 of a 4-chain run."""
    _need_gpu()
    X = datasets["K2_N100_P5"]
    K, ns = 2, 30
    rng = RRng(1)
    iz = np.stack([rng.sample_int(K, X.shape[0]) for _ in range(4)])
    a = B.gibbs_collapsed(X, ns, K, chains=4, seed=77, initial_K=iz)
    b = B.gibbs_collapsed(X, ns, K, chains=2, seed=77, initial_K=iz[2:], chain_offset=2)
    assert np.array_equal(a["z"][2:], b["z"])
    _close(a["alpha"][2:], b["alpha"], rtol=0)


def test_posterior_means_philox(oracle, datasets):
    """Independent Philox chains vs oracle chains: posterior means of theta and pi within MC error."""
    _need_gpu()
    X = datasets["K3_N1000_P5"]
    N, P = X.shape
    K, ns, burnin = 3, 400, 100
    g = B.gibbs_full(X, ns, K, burnin=burnin, relabel=True, burnrelabel=20, chains=16, seed=123)
    # label-invariant summaries: sorted pi, and theta rows ordered by pi
    def summary(pi, theta):  # pi S x K, theta K x P x S
        order = np.argsort(-pi.mean(0))
        return pi.mean(0)[order], theta.mean(2)[order]
    gs = [summary(g["pi"][c], g["theta_original"][c]) for c in range(16)]
    os_ = []
    for c in range(4):
        ip, th = _init_full(K, P, 100 + c)
        r = oracle.gibbs_full(X, ip, th, ns, K, burnin=burnin, seed=200 + c, probes=False)
        t = r.tail()
        os_.append(summary(t["pi"], t["theta"]))
    gpi = np.mean([s[0] for s in gs], 0); opi = np.mean([s[0] for s in os_], 0)
    gth = np.mean([s[1] for s in gs], 0); oth = np.mean([s[1] for s in os_], 0)
    sd_pi = np.std([s[0] for s in gs], 0) + 0.01
    sd_th = np.std([s[1] for s in gs], 0) + 0.02
    assert (np.abs(gpi - opi) < 4 * sd_pi).all(), (gpi, opi)
    assert (np.abs(gth - oth) < 4 * sd_th).all(), (gth, oth)
    # and the documented generating truth (R/bmm-mcmc.R:46-50) as a sanity band
    assert np.allclose(gpi, [0.6, 0.2, 0.2], atol=0.06)


def test_rejects_non_binary(datasets):
    _need_gpu()
    X = datasets["K2_N100_P5"].copy()
    X[3, 2] = 2
    with pytest.raises(_lib.BmmError) as e:
        B.gibbs_collapsed(X, 20, 2)
    assert e.value.code == -3


def test_dp_requires_symmetric_prior(datasets):
    _need_gpu()
    with pytest.raises(_lib.BmmError) as e:
        B.gibbs_dp(datasets["K2_N100_P5"], 20, beta=0.5, gamma=0.7)
    assert e.value.code == -4


# ---- grid path: one chain over the whole GPU (kern_big.cu) ----------------------------------------
@pytest.mark.parametrize("sampler", ["full", "stickbreaking"])
def test_grid_path_replay(oracle, datasets, sampler):
    """The grid-wide kernels consume the oracle's uniforms / parameter draws: z, counts bit-exact."""
    _need_gpu()
    X = datasets["K3_N1000_P5"]
    N, P = X.shape
    K, ns, burnin = (3, 40, 10) if sampler == "full" else (8, 30, 10)
    ip, th = _init_full(K, P, 5)
    of = oracle.gibbs_full if sampler == "full" else oracle.gibbs_stickbreaking
    gf = B.gibbs_full if sampler == "full" else B.gibbs_stickbreaking
    r = of(X, ip, th, ns, K, burnin=burnin, seed=11)
    g = gf(X, ns, K, burnin=burnin, initial_pi=ip, initial_theta=th, replay=_replay_of(r),
           probes=("probs", "loglik"), grid_path=True)
    t = r.tail()
    _close(g["loglik"][1:], r["loglik"][1:])
    _close(g["probs"][1:], r["probs"][1:], atol=1e-300)
    assert np.array_equal(g["z"], t["z"])
    _close(g["theta"], t["theta"], rtol=0)
    _close(g["pi"], t["pi"], rtol=0)
    _close(g["alpha"], t["alpha"], rtol=0)


@pytest.mark.parametrize("sampler,K", [("full", 3), ("stickbreaking", 6)])
def test_grid_path_equals_chain_path(datasets, sampler, K):
    """Same seed => the grid-wide kernels and the chain-per-block kernel produce the same chain, bit for bit
    (global-index Philox counters, same operation order)."""
    _need_gpu()
    X = datasets["K3_N1000_P5"]
    gf = B.gibbs_full if sampler == "full" else B.gibbs_stickbreaking
    a = gf(X, 60, K, seed=42)
    b = gf(X, 60, K, seed=42, grid_path=True)
    for k in ("z", "theta", "pi", "alpha"):
        assert np.array_equal(a[k], b[k]), k


def test_grid_path_fp32_close(oracle, datasets):
    """BMM_FP32 probability arithmetic: conditional probabilities within 1e-4 relative of the fp64 oracle."""
    _need_gpu()
    X = datasets["K3_N1000_P5"]
    N, P = X.shape
    K, ns = 3, 12
    ip, th = _init_full(K, P, 5)
    r = oracle.gibbs_full(X, ip, th, ns, K, burnin=2, seed=11)
    g = B.gibbs_full(X, ns, K, burnin=2, initial_pi=ip, initial_theta=th, replay=_replay_of(r), probes=("probs",),
                     grid_path=True, precision="fp32")
    _close(g["probs"][1:], r["probs"][1:], rtol=1e-4, atol=1e-30)


def test_grid_path_large_synthetic():
    """Full-size property checks (no oracle at this size): counts consistent with z, pi sums to 1."""
    _need_gpu()
    rng = np.random.default_rng(3)
    N, P, K = 200_000, 64, 32
    th_true = rng.uniform(0.1, 0.9, (K, P))
    zt = rng.integers(0, K, N)
    X = (rng.random((N, P)) < th_true[zt]).astype(np.int32)
    g = B.gibbs_stickbreaking(X, 12, K, alpha=1.0, burnin=1, seed=9)
    assert g["z"].shape == (11, N) and g["z"].min() >= 1 and g["z"].max() <= K
    assert np.allclose(g["pi"].sum(1), 1.0) and np.isfinite(g["theta"]).all()
    h = B.gibbs_stickbreaking(X, 12, K, alpha=1.0, burnin=1, seed=9, precision="fp32")
    assert h["z"].shape == (11, N)
    # first sweep uses identical parameters: fp32 and fp64 allocations agree except at ulp-close draws
    assert (g["z"][0] != h["z"][0]).mean() < 1e-3


# ---- tcgen05 sweep of the grid path (kern_big_ws.cu) ----------------------------------------------
def _host_counts(z, X, K):
    """[S][K + K*P]: c_k then V_kd (k + K*d) from an S x N allocation history."""
    S, P = z.shape[0], X.shape[1]
    out = np.zeros((S, K + K * P), dtype=np.int64)
    for s in range(S):
        oh = (z[s][:, None] == np.arange(1, K + 1)[None, :]).astype(np.int64)   # N x K
        out[s, :K] = oh.sum(0)
        out[s, K:] = (oh.T @ X).T.reshape(-1)                                   # (P, K) -> d-major
    return out


def test_grid_tensor_path_small(oracle, datasets):
    """tcgen05 log-likelihood contraction vs the fp64 oracle (1e-4 relative on the conditional
    probabilities) and tcgen05 sufficient statistics vs counts recomputed from the returned z (exact)."""
    _need_gpu()
    X = datasets["K3_N1000_P5"]
    N, P = X.shape
    K, ns = 3, 7
    ip, th = _init_full(K, P, 5)
    r = oracle.gibbs_full(X, ip, th, 2, K, burnin=0, seed=11)
    g = B.gibbs_full(X, ns, K, burnin=1, initial_pi=ip, initial_theta=th, seed=3, probes=("probs", "counts"),
                     grid_path=True, precision="fp32")
    _close(g["probs"][1], r["probs"][1], rtol=1e-4, atol=1e-30)
    assert np.array_equal(g["counts"][1:], _host_counts(g["z"], X, K))
    h = B.gibbs_full(X, ns, K, burnin=1, initial_pi=ip, initial_theta=th, seed=3, grid_path=True)   # fp64, same Philox
    assert (g["z"][0] != h["z"][0]).mean() < 2e-3


def test_grid_tensor_path_c4_shape():
    """C4 shape (P=64, K=32) with a ragged last tile: tensor-core path vs the CUDA-core float kernel."""
    _need_gpu()
    rng = np.random.default_rng(5)
    N, P, K = 50_000 + 37, 64, 32
    th_true = rng.uniform(0.1, 0.9, (K, P))
    X = (rng.random((N, P)) < th_true[rng.integers(0, K, N)]).astype(np.int32)
    kw = dict(alpha=1.0, burnin=1, seed=9, precision="fp32", probes=("probs", "counts"))
    g = B.gibbs_stickbreaking(X, 5, K, **kw)
    c = B.gibbs_stickbreaking(X, 5, K, no_tensor=True, **kw)
    assert np.array_equal(g["counts"][1:], _host_counts(g["z"], X, K))
    assert np.array_equal(c["counts"][1:], _host_counts(c["z"], X, K))
    big = c["probs"][1] > 1e-12
    _close(g["probs"][1][big], c["probs"][1][big], rtol=2e-4)
    assert (g["z"][0] != c["z"][0]).mean() < 2e-3
    assert np.allclose(g["pi"].sum(1), 1.0) and np.isfinite(g["theta"]).all()


@pytest.mark.parametrize("N,P,K", [(3077, 512, 100), (2000, 100, 50), (1500, 130, 40), (1000, 40, 64), (900, 33, 128)])
def test_grid_large_p_tensor_path(N, P, K):
    """Large-P / large-K tcgen05 path (kern_big_lp.cu) vs the CUDA-core fp64 grid kernel with the
    stabilised softmax (the reference's own exp underflows at large P): probabilities within 1e-4,
    counts exact, allocations equal except at ulp-close draws.  P not a multiple of 64 / 32 exercises
    the zero-padded last step and odd word counts."""
    _need_gpu()
    rng = np.random.default_rng(11)
    th_true = rng.uniform(0.2, 0.8, (K, P))
    X = (rng.random((N, P)) < th_true[rng.integers(0, K, N)]).astype(np.int32)
    ip = np.full(K, 1.0 / K)
    th0 = rng.uniform(0.3, 0.7, (K, P))
    kw = dict(alpha=1.0, burnin=1, seed=4, initial_pi=ip, initial_theta=th0, probes=("probs", "counts"), grid_path=True)
    g = B.gibbs_full(X, 4, K, precision="fp32", **kw)
    c = B.gibbs_full(X, 4, K, precision="fp64", stable_softmax=True, **kw)
    assert np.array_equal(g["counts"][1:], _host_counts(g["z"], X, K))
    big = c["probs"][1] > 1e-9
    _close(g["probs"][1][big], c["probs"][1][big], rtol=1e-4)
    assert (g["z"][0] != c["z"][0]).mean() < 5e-3
    assert np.allclose(g["pi"].sum(1), 1.0) and np.isfinite(g["theta"]).all()


def test_grid_large_p_tensor_path_vs_oracle(oracle):
    """The same path against the ORACLE itself (full_gibbs.cpp:92-122 with the max-subtraction the reference lacks,
    quirk 13) at P = 512, K = 100: conditional probabilities of the first sweep within 1e-4 relative (north_star's
    fp32 bar), for every entry that is not negligible."""
    _need_gpu()
    N, P, K = 3077, 512, 100
    rng = np.random.default_rng(11)
    th_true = rng.uniform(0.2, 0.8, (K, P))
    X = (rng.random((N, P)) < th_true[rng.integers(0, K, N)]).astype(np.int32)
    ip = np.full(K, 1.0 / K)
    th0 = rng.uniform(0.3, 0.7, (K, P))
    g = B.gibbs_full(X, 2, K, alpha=1.0, burnin=0, seed=4, initial_pi=ip, initial_theta=th0, probes=("probs",),
                     grid_path=True, precision="fp32")
    r = oracle.gibbs_full(X, ip, th0, 2, K, alpha=1.0, burnin=0, seed=4, stabilise=True)
    big = r["probs"][1] > 1e-9
    assert big.sum() > N
    _close(g["probs"][1][big], r["probs"][1][big], rtol=1e-4)
    assert np.abs(g["probs"][1] - r["probs"][1]).max() < 1e-4      # the negligible entries too, in absolute terms


def test_dp_philox_posterior(oracle, datasets):
    """gibbs_dp with Philox draws (inverse CDF over the used-list order) vs oracle chains (descending-sort
    walk of RcppArmadillo::sample): same posterior -- number of occupied clusters and co-clustering rate."""
    _need_gpu()
    X = datasets["K2_N100_P5"]
    ns, burnin = 300, 100

    def summary(z):  # z: S x N
        kact = np.mean([len(np.unique(r)) for r in z])
        co = np.mean(z[:, :10, None] == z[:, None, :10])     # co-clustering among the first ten observations
        return kact, co
    g = B.gibbs_dp(X, ns, burnin=burnin, maxK=30, chains=48, seed=31)
    gs = np.array([summary(g["z"][c]) for c in range(48)])
    os_ = np.array([summary(oracle.gibbs_dp(X, ns, alpha=0.0, burnin=burnin, maxK=30, seed=500 + c, probes=False).tail()["z"])
                    for c in range(12)])
    se = np.sqrt(gs.var(0) / 48 + os_.var(0) / 12)
    assert (np.abs(gs.mean(0) - os_.mean(0)) < 4 * se + 0.02).all(), (gs.mean(0), os_.mean(0), se)


def test_fetch_widening_equals_direct(datasets, monkeypatch):
    """Large int32 z outputs travel as bytes and are widened on the host: same arrays as the direct int32 download.
    With relabelling only z_original crosses PCIe and z = perm[z_original] (full_gibbs.cpp:171-174) is derived on the
    host while widening; BMM_FETCH_DERIVE=0 ships both matrices."""
    _need_gpu()
    X = datasets["K3_N1000_P5"]
    kw = dict(burnin=20, relabel=True, burnrelabel=10, chains=64, seed=8)
    a = B.gibbs_full(X, 200, 3, **kw)                    # 64 * 180 * 1000 = 11.5M allocations: staged + widened, z derived
    monkeypatch.setenv("BMM_FETCH_DERIVE", "0")
    c = B.gibbs_full(X, 200, 3, **kw)                    # both matrices as bytes
    monkeypatch.setenv("BMM_FETCH_WIDEN", "0")
    b = B.gibbs_full(X, 200, 3, **kw)                    # both matrices as int32
    for k in ("z", "z_original", "permutations", "theta", "pi"):
        assert np.array_equal(a[k], b[k]), k
        assert np.array_equal(c[k], b[k]), k
    assert a["z"].dtype == np.int32 and a["z"].min() >= 1 and a["z"].max() <= 3
    s_idx = np.arange(a["z"].shape[1])[None, :, None]
    c_idx = np.arange(64)[:, None, None]
    assert np.array_equal(a["z"], a["permutations"][c_idx, s_idx, a["z_original"] - 1] + 1)


@pytest.mark.parametrize("sampler", ["full", "collapsed", "dp"])
def test_fetch_derived_z_odd_shapes(datasets, monkeypatch, sampler):
    """The host-derived relabelled matrix for history lengths and chain counts that are not multiples of the vector
    width or the staging chunk (runs split across chunks and workers), with thinning and the z_freq summary."""
    _need_gpu()
    X = datasets["K3_N1000_P5"][:997]
    kw = dict(burnin=9, relabel=True, burnrelabel=5, chains=61, seed=5, thin=1, probes=("z_freq",))
    ns = 9 + 187
    run = {"full": lambda: B.gibbs_full(X, ns, 3, **kw), "collapsed": lambda: B.gibbs_collapsed(X, ns, 3, **kw),
           "dp": lambda: B.gibbs_dp(X, ns, maxK=12, **kw)}[sampler]
    a = run()                                            # 61 * 187 * 997 = 11.4M allocations
    monkeypatch.setenv("BMM_FETCH_WIDEN", "0")
    b = run()
    for k in ("z", "z_original", "permutations", "z_freq"):
        assert np.array_equal(a[k], b[k]), k
    K = a["permutations"].shape[-1]
    assert np.array_equal(a["z_freq"], np.stack([(a["z"] == k + 1).sum(1) for k in range(K)], axis=2))


def test_fetch_widening_above_16_labels(datasets, monkeypatch):
    """More than 16 labels: the relabelled matrix is not derived on the host (the byte table covers K <= 16), both matrices
    travel as bytes, still sweep segment by sweep segment."""
    _need_gpu()
    X = datasets["K2_N1000_P5"]
    kw = dict(burnin=20, relabel=True, burnrelabel=5, maxK=20, chains=40, seed=12)
    a = B.gibbs_dp(X, 20 + 300, **kw)                    # 40 * 300 * 1000 = 12M allocations, S = 300: one 256-slot segment + rest
    monkeypatch.setenv("BMM_FETCH_WIDEN", "0")
    b = B.gibbs_dp(X, 20 + 300, **kw)
    for k in ("z", "z_original", "permutations"):
        assert np.array_equal(a[k], b[k]), k


@pytest.mark.parametrize("name,K", [("K3_N1000_P5", 3), ("K2_N1000_P5", 2)])
def test_grid_path_relabel_replay(oracle, datasets, name, K):
    """Stephens batch + online relabelling with the streaming grid kernels (float P / Q, warp Jonker-Volgenant /
    enumeration) against the oracle (reference lp_solve): same permutations, same relabelled z, Q within fp32."""
    _need_gpu()
    X = datasets[name]
    N, P = X.shape
    ns, burnin, br = 50, 20, 6
    ip, th = _init_full(K, P, 5)
    r = oracle.gibbs_full(X, ip, th, ns, K, burnin=burnin, relabel=True, burnrelabel=br, seed=11)
    g = B.gibbs_full(X, ns, K, burnin=burnin, relabel=True, burnrelabel=br, initial_pi=ip, initial_theta=th,
                     replay=_replay_of(r), probes=("Q_final",), grid_path=True)
    t = r.tail()
    assert np.array_equal(g["z_original"], t["z_original"])
    assert np.array_equal(g["permutations"], t["permutations"])
    assert np.array_equal(g["z"], t["z"])
    _close(g["theta"], t["theta"], rtol=0)
    _close(g["theta_original"], t["theta_original"], rtol=0)
    _close(g["Q_final"], r["Q_final"], rtol=2e-4)


def test_grid_relabel_large_k_tensor_path():
    """Relabelling on the tcgen05 large-K path (K = 40 > enumeration range: warp Jonker-Volgenant): the
    permutation of every sweep is a permutation, z is z_original mapped through it, and the assignment is
    optimal for the cost matrix recomputed on the host from the same probabilities."""
    _need_gpu()
    rng = np.random.default_rng(2)
    N, P, K = 4000, 128, 40
    th_true = rng.uniform(0.15, 0.85, (K, P))
    X = (rng.random((N, P)) < th_true[rng.integers(0, K, N)]).astype(np.int32)
    burnin, M = 6, 2
    g = B.gibbs_full(X, 14, K, alpha=1.0, burnin=burnin, relabel=True, burnrelabel=M, seed=3, precision="fp32", probes=("probs",))
    S = g["permutations"].shape[0]
    assert np.array_equal(np.sort(g["permutations"], 1), np.tile(np.arange(K), (S, 1)))
    assert np.array_equal(g["z"], np.take_along_axis(g["permutations"], g["z_original"] - 1, 1) + 1)
    th_rel = np.empty_like(g["theta_original"])
    for s in range(S):
        th_rel[g["permutations"][s], :, s] = g["theta_original"][:, :, s]
    assert np.array_equal(th_rel, g["theta"])
    # optimality of the first online assignment: batch Q (stephens.cpp:6-64) and the online cost (:78-80) recomputed on
    # the host in float64 from the probabilities the kernel reported, solved by the reference's own lp_solve
    from oracle import pyoracle as O
    cube = np.stack([g["probs"][j] for j in range(burnin - M, burnin)], axis=2)       # N x K x M
    q, _ = O.stephens_batch(cube)
    pj = g["probs"][burnin]
    cost = np.array([[np.sum(pj[:, l] * (pj[:, l] - np.log(q[:, k]))) for l in range(K)] for k in range(K)])
    sol = O.assign(cost, O.has_ref())
    best = float((cost * sol).sum())
    perm = g["permutations"][0]                       # perm[l] = reference label matched to sample column l
    got = float(sum(cost[perm[l], l] for l in range(K)))
    assert got <= best + 1e-4 * abs(best), (got, best)


def _online_relabel_check(g, X, K, burnin, M, oracle, fixed=False):
    """Follow the relabelling of a grid-path run on the host in float64 from the probabilities the kernels reported: batch
    Q (stephens.cpp:6-64), then per kept sweep the online cost (:78-80) with the oracle's assignment (reference lp_solve)
    and the Q update (:87-92) along the permutation the GPU chose.  The GPU's permutation must be optimal for the host's
    cost at every sweep (identical when the optimum is unique), and its final Q must be the host's."""
    from oracle import pyoracle as O
    N = X.shape[0]
    ns = g["probs"].shape[0]
    cube = np.stack([g["probs"][j] for j in range(burnin - M, burnin)], axis=2)       # N x K x M
    q, _ = O.stephens_batch(cube)
    same = 0
    for j in range(burnin, ns):
        pj = g["probs"][j]
        perm_o, q_o, cost = O.stephens_online(q, pj, j, use_ref=O.has_ref())
        perm_g = g["permutations"][j - burnin]
        assert np.array_equal(np.sort(perm_g), np.arange(K))
        got = float(sum(cost[perm_g[l], l] for l in range(K)))
        best = float(sum(cost[perm_o[l], l] for l in range(K)))
        assert got <= best + 1e-7 * max(1.0, abs(best)), (j, got, best)
        same += int(np.array_equal(perm_g, perm_o))
        q = j * (q + pj[:, perm_g]) / (j + 1)                                         # quirks 3, 5
    assert same >= (ns - burnin) // 2          # ties aside, the very permutation of the oracle
    np.testing.assert_allclose(g["Q_final"], q, rtol=3e-4)
    assert np.array_equal(g["z"], np.take_along_axis(g["permutations"], g["z_original"] - 1, 1) + 1)


@pytest.mark.parametrize("sampler,N,P,K", [("full", 3000, 64, 8), ("stickbreaking", 20_000 + 77, 64, 32), ("full", 1111, 20, 5),
                                            ("full", 4096, 100, 24)])
def test_grid_tensor_relabel_single_pass(oracle, monkeypatch, sampler, N, P, K):
    """Online relabelling of the tensor path in one pass over Q per sweep (kern_big_ws_relabel.cu: probabilities recomputed
    on tcgen05 instead of stored, Q update deferred into the next sweep's pass, K x K contraction on tcgen05) against the
    oracle's Stephens step fed with the reported probabilities, and against the three-pass kernels it replaces."""
    _need_gpu()
    rng = np.random.default_rng(N + K)
    Kt = min(K, 8)
    th_true = rng.uniform(0.1, 0.9, (Kt, P))
    X = (rng.random((N, P)) < th_true[rng.integers(0, Kt, N)]).astype(np.int32)
    burnin, M, ns = 6, 3, 15
    run = B.gibbs_full if sampler == "full" else B.gibbs_stickbreaking
    kw = dict(alpha=1.0, burnin=burnin, relabel=True, burnrelabel=M, seed=5, precision="fp32", probes=("probs", "Q_final"),
              grid_path=True)
    g = run(X, ns, K, **kw)
    _online_relabel_check(g, X, K, burnin, M, oracle)
    monkeypatch.setenv("BMM_RELABEL_FUSED", "0")
    h = run(X, ns, K, **kw)                     # probabilities stored, cost kernel, Q update kernel
    assert np.array_equal(g["z_original"], h["z_original"])
    _online_relabel_check(h, X, K, burnin, M, oracle)
    np.testing.assert_allclose(g["Q_final"], h["Q_final"], rtol=3e-4)
    if np.array_equal(g["permutations"], h["permutations"]):
        assert np.array_equal(g["z"], h["z"]) and np.array_equal(g["theta"], h["theta"])


@pytest.mark.parametrize("N,P,K", [(20_000 + 77, 64, 32), (9_000 + 5, 100, 24), (12_000 + 3, 20, 7)])
def test_grid_tensor_kernels_many_tiles_per_cta(oracle, monkeypatch, N, P, K):
    """The warp-specialised tensor kernels (z-sweep and single-pass relabelling) with every mbarrier ring wrapping many
    times: the grid is capped at 3 CTAs, so each CTA walks 25-50 tiles (a full-size run has 528 per CTA; with the usual one
    or two tiles per CTA of a test-sized input no ring ever wraps).  P = 100 runs the two-warpgroup layout of the
    relabelling kernel, P = 20 the shallowest operand stages."""
    _need_gpu()
    rng = np.random.default_rng(99)
    th_true = rng.uniform(0.1, 0.9, (8, P))
    X = (rng.random((N, P)) < th_true[rng.integers(0, 8, N)]).astype(np.int32)
    burnin, M, ns = 6, 3, 13
    kw = dict(alpha=1.0, burnin=burnin, relabel=True, burnrelabel=M, seed=5, precision="fp32", probes=("probs", "Q_final", "counts"),
              grid_path=True)
    ref = B.gibbs_stickbreaking(X, ns, K, **kw)
    monkeypatch.setenv("BMM_GRID_MAX_CTAS", "3")
    g = B.gibbs_stickbreaking(X, ns, K, **kw)
    assert np.array_equal(g["z_original"], ref["z_original"])          # tile -> CTA mapping does not enter the draws
    assert np.array_equal(g["counts"][burnin:], _host_counts(g["z_original"], X, K))    # sufficient statistics of the kept sweeps
    _online_relabel_check(g, X, K, burnin, M, oracle)
    np.testing.assert_allclose(g["Q_final"], ref["Q_final"], rtol=3e-4)


def test_grid_tensor_relabel_single_pass_fixed_mode(monkeypatch):
    """The same single-pass kernel in the correctness-fixed Stephens mode (inverse permutation, running-mean Q): equal to
    the three-pass kernels."""
    _need_gpu()
    rng = np.random.default_rng(77)
    N, P, K = 5000 + 3, 64, 6
    th_true = rng.uniform(0.1, 0.9, (K, P))
    X = (rng.random((N, P)) < th_true[rng.integers(0, K, N)]).astype(np.int32)
    kw = dict(alpha=1.0, burnin=5, relabel=True, burnrelabel=3, seed=8, precision="fp32", probes=("Q_final",), grid_path=True,
              stephens_fixed=True)
    g = B.gibbs_full(X, 16, K, **kw)
    monkeypatch.setenv("BMM_RELABEL_FUSED", "0")
    h = B.gibbs_full(X, 16, K, **kw)
    assert np.array_equal(g["z_original"], h["z_original"])
    assert np.array_equal(g["permutations"], h["permutations"])
    assert np.array_equal(g["z"], h["z"])
    np.testing.assert_allclose(g["Q_final"], h["Q_final"], rtol=3e-4)


def test_assign_warp_jv_matches_lpsolve(oracle):
    """The warp-parallel Jonker-Volgenant solver of the grid path (through bmm_stephens-style cost) equals lp_solve."""
    _need_gpu()
    import ctypes as C
    L = _lib.lib()
    if not hasattr(L, "bmm_assign_warp"):
        pytest.skip("bmm_assign_warp not exported")
    rng = np.random.default_rng(4)
    for K in (6, 17, 32, 64, 128):
        cost = rng.uniform(0, 1000, (K, K))
        cf = np.asfortranarray(cost)
        perm = np.zeros(K, dtype=np.int32)
        _lib.check(L.bmm_assign_warp(K, cf.ctypes.data_as(C.POINTER(C.c_double)), perm.ctypes.data_as(C.POINTER(C.c_int32))))
        s_o = oracle.assign(cost, use_ref=oracle.has_ref())      # the reference's lp_solve, 0.25 s at K = 128
        assert np.array_equal(perm, s_o.argmax(0)), K


@pytest.mark.parametrize("N,K,use_logp", [(5000, 128, 0), (777, 72, 1), (64, 128, 0), (20011, 96, 1), (3001, 32, 0), (1000, 24, 1), (4097, 64, 0), (1003, 8, 0), (50001, 16, 1)])
def test_grid_cost_kernels_match_float64(N, K, use_logp):
    """Cost contraction of the grid path's relabelling (stephens.cpp:45-53,76-84): the tcgen05 kernel (fp16 hi/lo
    split operands) and the CUDA-core kernel against numpy float64, tolerance 2e-5 of the largest entry."""
    _need_gpu()
    import ctypes as C
    L = _lib.lib()
    rng = np.random.default_rng(N + K)
    p = rng.dirichlet(np.full(K, 0.3), N).astype(np.float32)
    q = rng.dirichlet(np.full(K, 0.5), N).astype(np.float32)
    q = np.maximum(q, np.float32(1e-6))
    p[p < 1e-30] = 0.0
    G = np.log(q.astype(np.float64)).T @ p.astype(np.float64)              # G[k, l]
    pd = p.astype(np.float64)
    s = np.where(pd > 0, pd * np.log(np.where(pd > 0, pd, 1.0)), 0.0).sum(0) if use_logp else (pd * pd).sum(0)
    fp = C.POINTER(C.c_float)
    for tensor in (1, 0):
        out = np.zeros(K * K + K)
        _lib.check(L.bmm_grid_cost(N, K, p.ctypes.data_as(fp), q.ctypes.data_as(fp), use_logp, tensor,
                                   out.ctypes.data_as(C.POINTER(C.c_double))))
        Gg = out[:K * K].reshape(K, K).T                                     # out[k + K*l]
        assert np.abs(Gg - G).max() <= 2e-5 * np.abs(G).max(), (tensor, np.abs(Gg - G).max(), np.abs(G).max())
        assert np.abs(out[K * K:] - s).max() <= 2e-5 * max(1.0, np.abs(s).max()), tensor


# ---- edge cases ----------------------------------------------------------------------------------
@pytest.mark.parametrize("N,P,K", [(1, 1, 1), (2, 1, 2), (7, 33, 3), (50, 64, 5), (40, 70, 9)])
def test_edge_shapes_replay_all_samplers(oracle, N, P, K):
    """Degenerate and ragged shapes (single observation / variable / cluster, P crossing a 32-bit word, K above
    the enumeration range) through every sampler, in replay against the oracle."""
    _need_gpu()
    rng = np.random.default_rng(N * 100 + P)
    X = (rng.random((N, P)) < 0.4).astype(np.int32)
    ns, burnin = 12, 3
    ip, th = _init_full(K, P, 3)
    r = oracle.gibbs_full(X, ip, th, ns, K, burnin=burnin, seed=2)
    for grid in (False, True):
        g = B.gibbs_full(X, ns, K, burnin=burnin, initial_pi=ip, initial_theta=th, replay=_replay_of(r), probes=("probs",),
                         grid_path=grid)
        _close(g["probs"][1:], r["probs"][1:], atol=1e-300)
        assert np.array_equal(g["z"], r.tail()["z"]), ("full", grid)
    r = oracle.gibbs_stickbreaking(X, ip, th, ns, K, burnin=burnin, seed=2)
    g = B.gibbs_stickbreaking(X, ns, K, burnin=burnin, initial_pi=ip, initial_theta=th, replay=_replay_of(r))
    assert np.array_equal(g["z"], r.tail()["z"])
    iz = RRng(3).sample_int(K, N)
    try:
        r = oracle.gibbs_collapsed(X, iz, ns, K, burnin=burnin, seed=2)
    except RuntimeError:
        # N = 1: removing the only observation empties every cluster, all probabilities are 0 and the
        # reference's rmultinom returns NA (collapsed_gibbs.cpp:104,131-133,154); the GPU flags it
        with pytest.raises(_lib.BmmError) as e:
            B.gibbs_collapsed(X, ns, K, burnin=burnin, initial_K=iz, seed=2)
        assert e.value.code == -9
        r = None
    if r is not None:
        g = B.gibbs_collapsed(X, ns, K, burnin=burnin, initial_K=iz, replay=_replay_of(r, keys=("alpha",)))
        assert np.array_equal(g["z"], r.tail()["z"])
    try:
        r = oracle.gibbs_dp(X, ns, alpha=0.0, burnin=burnin, maxK=max(K + 2, 4), seed=2)
    except RuntimeError:
        return
    g = B.gibbs_dp(X, ns, burnin=burnin, maxK=max(K + 2, 4), replay=_replay_of(r, keys=("alpha",)))
    assert np.array_equal(g["z"], r.tail()["z"])


def test_edge_burnin_zero_and_minimal_run(oracle, datasets):
    """burnin = 0 returns the initial state as iteration 0; nsamples = 2 is the shortest run."""
    _need_gpu()
    X = datasets["K2_N100_P5"]
    K = 2
    ip, th = _init_full(K, X.shape[1], 4)
    for grid in (False, True):
        g = B.gibbs_full(X, 2, K, burnin=0, initial_pi=ip, initial_theta=th, seed=1, grid_path=grid)
        assert g["theta"].shape == (K, 5, 2) and g["z"].shape == (2, 100)
        _close(g["theta"][:, :, 0], th, rtol=0)
        _close(g["pi"][0], ip, rtol=0)
        assert g["alpha"][0, 0] == 1.0
    iz = RRng(3).sample_int(K, 100)
    c = B.gibbs_collapsed(X, 2, K, burnin=0, initial_K=iz, seed=1)
    assert np.array_equal(c["z"][0], iz)


def test_edge_large_k_chain_path(oracle):
    """K = 72 clusters on the chain kernels: 4 labels per lane in the collapsed kernel, K x K cost matrix in global
    memory (above COST_SMEM_MAX), Hungarian relabelling."""
    _need_gpu()
    rng = np.random.default_rng(9)
    N, P, K = 300, 6, 72   # this seed also has a draw whose last two categories tie exactly (pp = 0.5)
    X = (rng.random((N, P)) < 0.5).astype(np.int32)
    ns, burnin, br = 10, 4, 2
    iz = RRng(3).sample_int(K, N)
    r = oracle.gibbs_collapsed(X, iz, ns, K, burnin=burnin, relabel=True, burnrelabel=br, seed=2, use_ref=False)
    g = B.gibbs_collapsed(X, ns, K, burnin=burnin, relabel=True, burnrelabel=br, initial_K=iz,
                          replay=_replay_of(r, keys=("alpha",)))
    assert np.array_equal(g["z_original"], r.tail()["z_original"])
    S = g["permutations"].shape[0]
    assert np.array_equal(np.sort(g["permutations"], 1), np.tile(np.arange(K), (S, 1)))


def test_invalid_arguments_are_rejected(datasets):
    _need_gpu()
    X = datasets["K2_N100_P5"]
    for kw, code in ((dict(nsamples=1, K=2), -1), (dict(nsamples=10, K=0), -1), (dict(nsamples=10, K=300), -1),
                     (dict(nsamples=10, K=2, burnin=10), -1), (dict(nsamples=10, K=2, beta=0.0), -1),
                     (dict(nsamples=20, K=2, burnin=1, relabel=True), -1)):
        with pytest.raises(_lib.BmmError) as e:
            B.gibbs_collapsed(X, kw.pop("nsamples"), kw.pop("K"), **kw)
        assert e.value.code == code, kw


def test_posterior_means_collapsed_and_stickbreaking(oracle, datasets):
    """Philox chains of the two remaining samplers against oracle chains: label-invariant posterior
    summaries (sorted cluster shares, theta rows ordered by share) within Monte Carlo error."""
    _need_gpu()
    X = datasets["K2_N1000_P5"]
    N, P = X.shape
    ns, burnin = 300, 100

    def summary(z, theta, K):                      # z: S x N, theta: K x P x S
        share = np.stack([(z == k + 1).mean(1) for k in range(K)], 1)            # S x K
        order = np.argsort(-share.mean(0))
        th = np.nanmean(theta, 2)[order]
        return share.mean(0)[order][:2], th[:2]

    # collapsed, K = 2
    g = B.gibbs_collapsed(X, ns, 2, burnin=burnin, chains=24, seed=5)
    gs = [summary(g["z"][c], g["theta"][c], 2) for c in range(24)]
    os_ = []
    for c in range(6):
        iz = RRng(40 + c).sample_int(2, N)
        t = oracle.gibbs_collapsed(X, iz, ns, 2, burnin=burnin, seed=300 + c, probes=False).tail()
        os_.append(summary(t["z"], t["theta"], 2))
    for idx in (0, 1):
        gm, om = np.mean([s[idx] for s in gs], 0), np.mean([s[idx] for s in os_], 0)
        se = np.std([s[idx] for s in gs], 0) / np.sqrt(24) + np.std([s[idx] for s in os_], 0) / np.sqrt(6) + 0.01
        assert (np.abs(gm - om) < 4 * se).all(), ("collapsed", idx, gm, om)
    assert np.allclose(np.mean([s[0] for s in gs], 0), [0.7, 0.3], atol=0.05)     # documented truth, R/bmm-mcmc.R:31-35

    # stick-breaking, maxK = 6: the two occupied sticks carry the same shares
    g = B.gibbs_stickbreaking(X, ns, 6, burnin=burnin, chains=24, seed=7)
    gs = [summary(g["z"][c], g["theta"][c], 6) for c in range(24)]
    os_ = []
    for c in range(4):
        ip, th = _init_full(6, P, 70 + c)
        t = oracle.gibbs_stickbreaking(X, ip, th, ns, 6, burnin=burnin, seed=400 + c, probes=False).tail()
        os_.append(summary(t["z"], t["theta"], 6))
    for idx in (0, 1):
        gm, om = np.mean([s[idx] for s in gs], 0), np.mean([s[idx] for s in os_], 0)
        se = np.std([s[idx] for s in gs], 0) / np.sqrt(24) + np.std([s[idx] for s in os_], 0) / np.sqrt(4) + 0.015
        assert (np.abs(gm - om) < 4 * se).all(), ("stickbreaking", idx, gm, om)


@pytest.mark.parametrize("relabel,precision", [(False, "fp64"), (True, "fp64"), (True, "fp32")])
def test_grid_posterior_summaries(datasets, relabel, precision):
    """z_freq / z_last (what a large-N caller keeps instead of the S x N history) equal the same summaries
    computed from the full history; and they are available with no_z_history.  fp32 = the tensor path, whose
    relabelling runs its assignment on a side stream: the summary must wait for it."""
    _need_gpu()
    X = datasets["K3_N1000_P5"]
    K, ns, burnin = 3, 60, 20
    kw = dict(burnin=burnin, relabel=relabel, burnrelabel=5, seed=6, grid_path=True, precision=precision)
    g = B.gibbs_full(X, ns, K, probes=("z_freq", "z_last"), **kw)
    freq = np.stack([(g["z"] == k + 1).sum(0) for k in range(K)], 1)
    assert np.array_equal(g["z_freq"], freq)
    assert np.array_equal(g["z_last"], (g["z_original"] if relabel else g["z"])[-1])
    h = B.gibbs_full(X, ns, K, probes=("z_freq", "z_last"), no_z_history=True, **kw)
    assert "z" not in h and np.array_equal(h["z_freq"], g["z_freq"]) and np.array_equal(h["z_last"], g["z_last"])
    _close(h["theta"], g["theta"], rtol=0)


@pytest.mark.parametrize("sampler", ["full", "stickbreaking", "collapsed", "dp", "grid", "grid_tensor"])
def test_thinning_keeps_every_kth_sweep(datasets, sampler):
    """thin = k (f2): every returned history holds the sweeps burnin, burnin + k, ... of the un-thinned run -- same
    chain, same relabelling (which still runs every sweep), a k-th of the storage."""
    _need_gpu()
    X = datasets["K2_N100_P5"]
    ns, burnin, thin = 48, 11, 3
    kw = dict(burnin=burnin, relabel=True, burnrelabel=4, seed=13)
    if sampler in ("full", "stickbreaking", "grid", "grid_tensor"):
        f = B.gibbs_stickbreaking if sampler == "stickbreaking" else B.gibbs_full
        kw.update(chains=1 if sampler.startswith("grid") else 2, grid_path=sampler.startswith("grid"),
                  precision="fp32" if sampler == "grid_tensor" else "fp64")
        run = lambda **e: f(X, ns, 4, **kw, **e)
    elif sampler == "collapsed":
        kw.update(chains=2)
        run = lambda **e: B.gibbs_collapsed(X, ns, 3, **kw, **e)
    else:
        kw.update(chains=2)
        run = lambda **e: B.gibbs_dp(X, ns, maxK=12, **kw, **e)
    full = run()
    thinned = run(thin=thin)
    S = ns - burnin
    keep = np.arange(0, S, thin)
    assert set(full) == set(thinned)
    for k, v in full.items():
        multi = kw["chains"] > 1
        if k in ("theta", "theta_original"):
            want = v[..., keep]
        elif multi:
            want = v[:, keep]
        else:
            want = v[keep]
        assert thinned[k].shape == want.shape, (k, thinned[k].shape, want.shape)
        assert np.array_equal(thinned[k], want, equal_nan=True), (sampler, k)


@pytest.mark.parametrize("sampler,relabel", [("full", True), ("collapsed", False), ("dp", True)])
def test_chain_path_posterior_summaries(datasets, sampler, relabel):
    """z_freq / z_last on the chain-parallel paths (f2): allocation counts over the kept sweeps (relabelled when
    relabel) and the last sweep's allocations, equal to the same summaries computed from the returned history."""
    _need_gpu()
    X = datasets["K2_N100_P5"]
    N = X.shape[0]
    ns, burnin, C_ = 40, 10, 3
    kw = dict(burnin=burnin, relabel=relabel, burnrelabel=4, seed=3, chains=C_, probes=("z_freq", "z_last"), thin=2)
    if sampler == "full":
        K = 3
        g = B.gibbs_full(X, ns, K, **kw)
    elif sampler == "collapsed":
        K = 3
        g = B.gibbs_collapsed(X, ns, K, **kw)
    else:
        K = 10
        g = B.gibbs_dp(X, ns, maxK=K, **kw)
    assert g["z_freq"].shape == (C_, N, K) and g["z_last"].shape == (C_, N)
    freq = np.stack([(g["z"] == k + 1).sum(1) for k in range(K)], axis=2)
    assert np.array_equal(g["z_freq"], freq)
    zo = g["z_original"] if relabel else g["z"]
    if (ns - 1 - burnin) % 2 == 0:      # the last sweep is a kept one
        assert np.array_equal(g["z_last"], zo[:, -1])
    assert g["z_last"].min() >= 1 and g["z_last"].max() <= K


def test_predictive_distribution(datasets):
    """Posterior predictive distribution from the kept draws (the reference only lists it in its TODO file, so
    the check is the definition itself in float64): log 1/S sum_s sum_k pi_k prod_d theta^x (1 - theta)^(1 - x) and the
    draw-averaged responsibilities, single chain and pooled chains, P above one packed word."""
    _need_gpu()
    X = datasets["K3_N1000_P5"]
    fit = B.gibbs_full(X, 120, 3, burnin=20, relabel=True, burnrelabel=5, seed=4)
    new = np.array([[a, b, c, d, e] for a in (0, 1) for b in (0, 1) for c in (0, 1) for d in (0, 1) for e in (0, 1)], dtype=np.int32)

    def host(theta, pi, Xn):
        K, P, S = theta.shape
        ll = (Xn[:, None, :, None] * np.log(theta)[None] + (1 - Xn)[:, None, :, None] * np.log1p(-theta)[None]).sum(2)   # M x K x S
        lw = ll + np.log(pi.T)[None]                                                                                    # M x K x S
        mx = lw.max(1, keepdims=True)
        per_draw = mx[:, 0] + np.log(np.exp(lw - mx).sum(1))                                                            # M x S
        resp = np.exp(lw - per_draw[:, None, :]).mean(2)
        m2 = per_draw.max(1, keepdims=True)
        return m2[:, 0] + np.log(np.exp(per_draw - m2).mean(1)), resp

    g = B.predictive(fit, new)
    lp, resp = host(fit["theta"], fit["pi"], new)
    np.testing.assert_allclose(g["log_pred"], lp, rtol=1e-10)
    np.testing.assert_allclose(g["membership"], resp, rtol=1e-9, atol=1e-14)
    assert abs(np.exp(g["log_pred"]).sum() - 1.0) < 1e-9            # the 32 patterns of P = 5 exhaust the sample space
    np.testing.assert_allclose(g["membership"].sum(1), 1.0, rtol=1e-12)
    # pooled chains
    fit4 = B.gibbs_full(X, 60, 3, burnin=10, chains=4, seed=9)
    g4 = B.predictive(fit4, new, membership=False)
    lp4, _ = host(np.concatenate(list(fit4["theta"]), axis=2), np.concatenate(list(fit4["pi"]), axis=0), new)
    np.testing.assert_allclose(g4["log_pred"], lp4, rtol=1e-10)
    # P = 70 (three packed words), stick-breaking
    rng = np.random.default_rng(1)
    th = rng.uniform(0.1, 0.9, (4, 70))
    Xb = (rng.random((600, 70)) < th[rng.integers(0, 4, 600)]).astype(np.int32)
    fb = B.gibbs_stickbreaking(Xb, 40, 6, alpha=1.0, burnin=10, seed=2)
    gb = B.predictive(fb, Xb[:50])
    lpb, respb = host(fb["theta"], fb["pi"], Xb[:50])
    np.testing.assert_allclose(gb["log_pred"], lpb, rtol=1e-10)
    np.testing.assert_allclose(gb["membership"], respb, rtol=1e-8, atol=1e-14)
    with pytest.raises(ValueError):
        B.predictive(B.gibbs_collapsed(X, 30, 3, seed=1), new)
