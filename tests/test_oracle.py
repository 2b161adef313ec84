"""The oracle against what pins it and against itself (invariants).

Pins: (1) reference lp_solve (oracle/_ref) for the assignment step; (2) the RNG KATs in
test_fixtures.py; (3) the documented generating truth of the bundled data as a sanity band;
(4) committed golden outputs (tests/golden/oracle_golden.npz, made by tools/make_golden.py) so a
change of the restatement cannot go unnoticed.  The samplers themselves are "parity unpinned":
the reference has no tests and R is not available here.
"""
import itertools
import os

import numpy as np
import pytest

from bmm_mcmc_b200.rcompat import RRng
from conftest import ROOT


def _init_full(K, P, seed):
    rng = RRng(seed)
    ip = np.exp(rng.runif(K)); ip /= ip.sum()
    return ip, rng.runif(K * P).reshape(P, K).T


def test_assign_hungarian_equals_reference_lpsolve(oracle):
    if not oracle.has_ref():
        pytest.skip("oracle/_ref/liblpsolve_ref.so not built (no /root/reference)")
    rng = np.random.default_rng(0)
    for K in (2, 3, 5, 8, 16):
        for _ in range(10):
            c = rng.uniform(0, 1000, (K, K))
            a, b = oracle.assign(c, True), oracle.assign(c, False)
            assert np.array_equal(a, b)
            assert (a.sum(0) == 1).all() and (a.sum(1) == 1).all()


def test_assign_is_optimal_by_enumeration(oracle):
    rng = np.random.default_rng(1)
    for K in (2, 3, 4, 5):
        c = rng.normal(size=(K, K))
        sol = oracle.assign(c, oracle.has_ref())
        best = min(sum(c[r, p[r]] for r in range(K)) for p in itertools.permutations(range(K)))
        assert np.isclose((c * sol).sum(), best)


def test_lpsolve_all_zero_cost_tiebreak(oracle):
    # SURVEY 8a11 [probe]: an all-zero cost gives the anti-diagonal in the reference solver
    if not oracle.has_ref():
        pytest.skip("no reference lp_solve")
    sol = oracle.assign(np.zeros((3, 3)), True)
    assert np.array_equal(sol, np.eye(3, dtype=np.int32)[::-1])


def test_rmultinom_rule(oracle):
    # zero categories are skipped without consuming a uniform; at most K-1 uniforms; one-hot result
    rc, rN, u = oracle.rmultinom1(3, [0.0, 0.3, 0.0, 0.7])
    assert rc == 0 and rN.sum() == 1 and rN[0] == 0 and rN[2] == 0 and len(u) <= 1
    rc, rN, u = oracle.rmultinom1(3, [0.2, 0.3, 0.5])
    assert rc == 0 and rN.sum() == 1 and 1 <= len(u) <= 2
    rc, _, _ = oracle.rmultinom1(3, [0.2, np.nan, 0.5])
    assert rc != 0
    # empirical frequencies
    cnt = np.zeros(3)
    for s in range(3000):
        cnt += oracle.rmultinom1(s, [0.2, 0.3, 0.5])[1]
    np.testing.assert_allclose(cnt / 3000, [0.2, 0.3, 0.5], atol=0.03)


def test_gamma_beta_moments(oracle):
    g = oracle.rgamma(1, 40000, 2.5, 2.0)
    assert abs(g.mean() - 5.0) < 0.08 and abs(g.var() - 10.0) < 0.5
    g = oracle.rgamma(2, 40000, 0.5)
    assert abs(g.mean() - 0.5) < 0.02
    b = oracle.rbeta(3, 40000, 0.5, 3.5)
    assert abs(b.mean() - 0.125) < 0.005
    d = oracle.rdirichlet(4, [1.0, 2.0, 3.0])
    assert np.isclose(d.sum(), 1.0) and (d > 0).all()


def test_stephens_online_formula(oracle):
    # Q' = j*(Q + p[:, perm])/(j+1) and cost = sum p*(p - log q)  (stephens.cpp:79,92)
    rng = np.random.default_rng(2)
    N, K, j = 50, 3, 9
    q = rng.uniform(0.1, 1, (N, K)); p = rng.dirichlet(np.ones(K), N)
    perm, qn, cost = oracle.stephens_online(q, p, j, use_ref=oracle.has_ref())
    want = np.array([[np.sum(p[:, l] * (p[:, l] - np.log(q[:, k]))) for l in range(K)] for k in range(K)])
    np.testing.assert_allclose(cost, want, rtol=1e-12)
    np.testing.assert_allclose(qn, j * (q + p[:, perm]) / (j + 1), rtol=1e-14)
    sol = oracle.assign(cost, oracle.has_ref())
    assert np.array_equal(perm, sol.argmax(0))


def test_stephens_batch_identity_when_consistent(oracle):
    # identically-labelled slices: the identity permutation is optimal and q is the slice mean
    rng = np.random.default_rng(3)
    N, K, M = 80, 3, 6
    base = rng.dirichlet([8, 1, 1], N)
    base[N // 3:2 * N // 3] = rng.dirichlet([1, 8, 1], N // 3 + (2 * N // 3 - N // 3 - N // 3))[: 2 * N // 3 - N // 3]
    base[2 * N // 3:] = rng.dirichlet([1, 1, 8], N - 2 * N // 3)
    p = np.stack([base] * M, axis=2)
    q, perm = oracle.stephens_batch(p, use_ref=oracle.has_ref())
    assert np.array_equal(perm, np.tile(np.arange(K), (M, 1)))
    np.testing.assert_allclose(q, base, rtol=1e-12)


def test_stephens_fixed_mode_undoes_a_three_cycle(oracle):
    """Correctness-fixed mode: with the sample's columns a 3-cycle of the reference's, the online
    step returns that cycle, re-orders with its inverse and keeps Q a running mean; the reference's own step
    (quirks 3-5) re-orders with the forward permutation, which is wrong for a 3-cycle."""
    rng = np.random.default_rng(5)
    N, K, j = 60, 3, 7
    base = rng.dirichlet(np.ones(K) * 0.3, N)
    sigma = np.array([1, 2, 0])
    p = base[:, sigma]                                   # sample column l shows reference label sigma[l]
    try:
        oracle.set_stephens_fixed(True)
        perm, qn, cost = oracle.stephens_online(base, p, j, use_ref=oracle.has_ref())
        want = np.array([[np.sum(p[:, l] * (np.log(p[:, l]) - np.log(base[:, k]))) for l in range(K)] for k in range(K)])
        np.testing.assert_allclose(cost, want, rtol=1e-12)
        assert np.array_equal(perm, sigma)
        np.testing.assert_allclose(qn, (j * base + base) / (j + 1), rtol=1e-14)     # running mean of consistent labels
        # batch: slices that are column permutations of one matrix come back to it
        sig = [np.array([0, 1, 2]), sigma, np.array([2, 0, 1]), np.array([0, 1, 2])]
        cube = np.stack([base[:, sg] for sg in sig], axis=2)
        q, inv = oracle.stephens_batch(cube, use_ref=oracle.has_ref())
        for t, sg in enumerate(sig):
            assert np.array_equal(inv[t], np.argsort(sg)), t      # reference label -> sample column
        np.testing.assert_allclose(q, np.maximum(base, 0), rtol=1e-12)
    finally:
        oracle.set_stephens_fixed(False)
    perm_q, qn_q, _ = oracle.stephens_online(base, p, j, use_ref=oracle.has_ref())
    assert np.array_equal(perm_q, sigma)
    assert not np.allclose(qn_q, j * (base + base) / (j + 1))       # forward re-ordering: columns stay mixed up


def test_full_posterior_recovers_truth(oracle, datasets):
    X = datasets["K3_N1000_P5"]
    ip, th = _init_full(3, 5, 1)
    r = oracle.gibbs_full(X, ip, th, 400, 3, burnin=150, seed=3, probes=False)
    t = r.tail()
    order = np.argsort(-t["pi"].mean(0))
    np.testing.assert_allclose(t["pi"].mean(0)[order], [0.6, 0.2, 0.2], atol=0.06)   # R/bmm-mcmc.R:46-50
    assert abs(t["theta"].mean(2)[order[0], 0] - 0.7) < 0.06


def test_collapsed_invariants(oracle, datasets):
    X = datasets["K2_N100_P5"]
    iz = RRng(1).sample_int(2, 100)
    r = oracle.gibbs_collapsed(X, iz, 60, 2, burnin=10, relabel=True, burnrelabel=5, seed=2)
    assert np.array_equal(r["z"][0], iz)                       # z_out.row(0) = initialK
    np.testing.assert_allclose(r["probs"][1:].sum(2), 1.0, rtol=1e-12)
    t = r.tail()
    # theta point estimate = S_kd / N_k of that sweep's allocations
    z = t["z_original"][-1]
    for k in range(2):
        np.testing.assert_allclose(t["theta_original"][k, :, -1], X[z == k + 1].mean(0))
    # relabelled z is the permutation of the original
    s = 7
    assert np.array_equal(t["z"][s], t["permutations"][s][t["z_original"][s] - 1] + 1)


def test_dp_invariants(oracle, datasets):
    X = datasets["K2_N1000_P5"]
    r = oracle.gibbs_dp(X, 60, maxK=64, seed=4)
    assert r["Kactive"][1:].min() >= 1 and r["Kactive"].max() < 63
    z = r["z"][1:]
    assert z.min() >= 1 and z.max() <= 64
    np.testing.assert_allclose(r["probs"][2:].sum(2), 1.0, rtol=1e-9)
    with pytest.raises(RuntimeError):
        oracle.gibbs_dp(X, 10, beta=0.5, gamma=0.6)            # Rcpp::stop (collapsed_gibbs_dp.cpp:48-50)


def test_golden_outputs_unchanged(oracle, datasets):
    path = os.path.join(ROOT, "tests", "golden", "oracle_golden.npz")
    g = np.load(path)
    from tools.make_golden import cases
    for name, fn in cases(oracle, datasets).items():
        out = fn()
        for key, val in out.items():
            ref = g["%s__%s" % (name, key)]
            if np.issubdtype(ref.dtype, np.integer):
                assert np.array_equal(val, ref), (name, key)
            else:
                np.testing.assert_allclose(val, ref, rtol=1e-12, atol=0, equal_nan=True, err_msg="%s %s" % (name, key))


def test_optimised_cpu_baseline_runs(datasets):
    """bench.py's extra.cpu_optimised leg (oracle/opt_cpu.cpp): builds, runs a few chains on two threads, rejects shapes it
    does not cover.  It is only ever timed, never compared."""
    from oracle import pyoracle as O
    X = datasets["K3_N1000_P5"]
    assert O.opt_cpu_full_gibbs(X, 3, 60, 10, 4, 2) > 0.0
    with pytest.raises(RuntimeError):
        O.opt_cpu_full_gibbs(np.zeros((10, 70), dtype=np.int32), 3, 10, 2, 1, 1)
    assert O.opt_cpu_collapsed_gibbs(X, 3, 40, 4, 2) > 0.0
