"""Host-side mirror of the reference's R user API (R/utils.R:23-107) over the C ABI.

Same function names, argument names, defaults and returned list layout as the R package:
`gibbs_full`, `gibbs_collapsed`, `gibbs_dp`, `gibbs_stickbreaking`.  The returned dict has the
R list's names and shapes (`pi` S x K, `alpha` S x 1, `permutations` S x K, `z` S x N,
`theta` K x P x S, and with `relabel=True` also `z_original`, `theta_original`); labels in `z`
are 1-based like R's.  Extra keyword-only arguments (`chains`, `seed`, `device`, ...) default to the
reference's behaviour; with `chains > 1` every array gains a leading chain axis.

The initial state is drawn on the host like the R wrappers do (R/utils.R:42,68-74,98-103), with an
R-compatible Mersenne-Twister (`rcompat.RRng`) so `seed=s` plays the role of `set.seed(s)`.
All sampling runs in the CUDA library; there is no CPU fallback.
"""
import ctypes as C
import numpy as np

from . import _lib
from .rcompat import RRng

__all__ = ["gibbs_full", "gibbs_collapsed", "gibbs_dp", "gibbs_stickbreaking"]


def _r_round(x):
    return int(round(x))  # R's round() and Python's are both half-to-even


def _as_X(data):
    X = np.asarray(data)
    if X.ndim != 2:
        raise ValueError("data must be a 2-d matrix (observations x binary variables)")
    if not np.issubdtype(X.dtype, np.integer):
        Xi = X.astype(np.int32)
        if not np.array_equal(Xi, X):
            raise ValueError("data must be integer 0/1")
        X = Xi
    return np.asfortranarray(X, dtype=np.int32)


def _p(arr, typ):
    return arr.ctypes.data_as(C.POINTER(typ)) if arr is not None else None


def _chain_cm(C_, shape, dtype):
    """Array of shape (C_, *shape) whose per-chain blocks are column-major (R layout)."""
    rev = tuple(reversed(shape))
    base = np.zeros((C_,) + rev, dtype=dtype)
    axes = (0,) + tuple(range(len(shape), 0, -1))
    return base.transpose(axes)


def _run(sampler, X, nsamples, K, alpha, beta, gamma, a, b, burnin, relabel, burnrelabel, debug,
         chains, seed, device, precision, init_pi=None, init_theta=None, init_z=None, replay=None,
         compact_z=False, stable_softmax=False, probes=(), chain_offset=0):
    L = _lib.lib()
    N, P = X.shape
    S = nsamples - burnin
    Cn = int(chains)
    args = _lib.Args()
    args.X = _p(X, C.c_int32)
    args.N, args.P, args.nsamples, args.K = N, P, int(nsamples), int(K)
    args.alpha, args.beta, args.gamma, args.a, args.b = float(alpha), float(beta), float(gamma), float(a), float(b)
    args.burnin, args.relabel, args.burnrelabel, args.debug = int(burnin), int(bool(relabel)), int(burnrelabel), int(bool(debug))
    args.n_chains, args.chain_offset, args.seed = Cn, int(chain_offset), int(seed) & 0xFFFFFFFFFFFFFFFF
    args.precision = {"fp64": _lib.BMM_FP64, "fp32": _lib.BMM_FP32}[precision]
    args.device = int(device)
    args.flags = (_lib.FLAG_COMPACT_Z if compact_z else 0) | (_lib.FLAG_STABLE_SOFTMAX if stable_softmax else 0)
    keep = []
    if replay is not None:
        rp = _lib.Replay()
        u = np.ascontiguousarray(replay["u"], dtype=np.float64)
        rp.u = _p(u, C.c_double)
        rp.u_slots = int(u.shape[-1])
        keep.append(u)
        for name in ("pi", "theta", "alpha"):
            if replay.get(name) is not None:
                arr = replay[name]
                keep.append(arr)
                setattr(rp, name, _p(arr, C.c_double))
        keep.append(rp)
        args.replay = C.pointer(rp)
    init = _lib.Init()
    if init_pi is not None:
        init.pi = _p(init_pi, C.c_double)
        init.theta = _p(init_theta, C.c_double)
    if init_z is not None:
        init.z = _p(init_z, C.c_int32)
    zt = np.uint8 if compact_z else np.int32
    res = {}
    out = _lib.Out()
    has_pi = sampler in (_lib.SAMPLER_FULL, _lib.SAMPLER_STICKBREAKING)
    if has_pi:
        res["pi"] = _chain_cm(Cn, (S, K), np.float64)
        out.pi = _p(res["pi"], C.c_double)
    res["alpha"] = _chain_cm(Cn, (S, 1), np.float64)
    out.alpha = _p(res["alpha"], C.c_double)
    res["permutations"] = _chain_cm(Cn, (S, K), np.int32)
    out.permutations = _p(res["permutations"], C.c_int32)
    res["z"] = _chain_cm(Cn, (S, N), zt)
    out.z = C.cast(res["z"].ctypes.data, C.POINTER(C.c_int32))
    res["theta"] = _chain_cm(Cn, (K, P, S), np.float64)
    out.theta = _p(res["theta"], C.c_double)
    if relabel:
        res["z_original"] = _chain_cm(Cn, (S, N), zt)
        out.z_original = C.cast(res["z_original"].ctypes.data, C.POINTER(C.c_int32))
        res["theta_original"] = _chain_cm(Cn, (K, P, S), np.float64)
        out.theta_original = _p(res["theta_original"], C.c_double)
    extra = {}
    if "probs" in probes:
        extra["probs"] = np.zeros((Cn, nsamples, K, N)).transpose(0, 1, 3, 2)  # [c][j] blocks of N x K cm
        out.probs = C.cast(extra["probs"].ctypes.data, C.POINTER(C.c_double))
    if "loglik" in probes:
        extra["loglik"] = np.zeros((Cn, nsamples, K, N)).transpose(0, 1, 3, 2)
        out.loglik = C.cast(extra["loglik"].ctypes.data, C.POINTER(C.c_double))
    if "Q_final" in probes and relabel:
        extra["Q_final"] = _chain_cm(Cn, (N, K), np.float64)
        out.Q_final = _p(extra["Q_final"], C.c_double)
    status = np.zeros(Cn, dtype=np.int32)
    out.status = _p(status, C.c_int32)
    fn = {_lib.SAMPLER_FULL: L.bmm_gibbs_full, _lib.SAMPLER_STICKBREAKING: L.bmm_gibbs_stickbreaking,
          _lib.SAMPLER_COLLAPSED: L.bmm_gibbs_collapsed}.get(sampler)
    if sampler == _lib.SAMPLER_DP:
        rc = L.bmm_gibbs_dp(C.byref(args), C.byref(out))
    else:
        rc = fn(C.byref(args), C.byref(init), C.byref(out))
    _lib.check(rc)
    res.update(extra)
    if Cn == 1 and chains == 1:
        res = {k: v[0] for k, v in res.items()}
    return res


def _defaults(nsamples, burnin, burnrelabel, alpha, clamp=True):
    if burnin is None:
        burnin = _r_round(0.1 * nsamples)
    if clamp and burnrelabel > burnin:
        burnrelabel = _r_round(0.1 * burnin)
    if alpha is None:
        alpha = 0
    return burnin, burnrelabel, alpha


def gibbs_full(data, nsamples, K, alpha=None, beta=0.5, gamma=0.5, a=1, b=1, burnin=None, relabel=False,
               burnrelabel=50, debug=False, *, chains=1, seed=0, device=0, precision="fp64", rng=None,
               initial_pi=None, initial_theta=None, replay=None, compact_z=False, stable_softmax=False,
               probes=(), chain_offset=0, _sampler=None):
    """Full Gibbs sampler for a finite Bernoulli mixture model (R/utils.R:64-78 -> full_gibbs.cpp:32)."""
    X = _as_X(data)
    N, P = X.shape
    burnin, burnrelabel, alpha = _defaults(nsamples, burnin, burnrelabel, alpha, clamp=_sampler is None)
    rng = rng or RRng(seed)
    if initial_pi is None:
        initial_pi = np.empty((chains, K))
        initial_theta = np.empty((chains, P, K))
        for c in range(chains):
            ip = np.exp(rng.runif(K))                       # R/utils.R:68-70
            initial_pi[c] = ip / ip.sum()
            initial_theta[c] = rng.runif(K * P).reshape(P, K)  # matrix(runif(K*P), nrow=K): column-major
    else:
        initial_pi = np.ascontiguousarray(np.asarray(initial_pi, dtype=np.float64).reshape(chains, K))
        th = np.asarray(initial_theta, dtype=np.float64).reshape(chains, K, P)
        initial_theta = np.ascontiguousarray(th.transpose(0, 2, 1))
    sampler = _lib.SAMPLER_FULL if _sampler is None else _sampler
    return _run(sampler, X, nsamples, K, alpha, beta, gamma, a, b, burnin, relabel, burnrelabel, debug, chains, seed,
                device, precision, init_pi=np.ascontiguousarray(initial_pi), init_theta=np.ascontiguousarray(initial_theta),
                replay=replay, compact_z=compact_z, stable_softmax=stable_softmax, probes=probes, chain_offset=chain_offset)


def gibbs_stickbreaking(data, nsamples, maxK, alpha=None, beta=0.5, gamma=0.5, a=1, b=1, burnin=None,
                        relabel=False, burnrelabel=50, debug=False, **kw):
    """Stick-breaking blocked Gibbs sampler (R/utils.R:95-107 -> stickbreaking.cpp:10).

    Like the reference wrapper it does NOT clamp burnrelabel to burnin (R/utils.R:97-101)."""
    return gibbs_full(data, nsamples, maxK, alpha, beta, gamma, a, b, burnin, relabel, burnrelabel, debug,
                      _sampler=_lib.SAMPLER_STICKBREAKING, **kw)


def gibbs_collapsed(data, nsamples, K, alpha=None, beta=0.5, gamma=0.5, a=1, b=1, burnin=None, relabel=False,
                    burnrelabel=50, debug=False, *, chains=1, seed=0, device=0, precision="fp64", rng=None,
                    initial_K=None, replay=None, compact_z=False, probes=(), chain_offset=0):
    """Collapsed Gibbs sampler for a finite mixture (R/utils.R:37-47 -> collapsed_gibbs.cpp:24)."""
    X = _as_X(data)
    N, P = X.shape
    burnin, burnrelabel, alpha = _defaults(nsamples, burnin, burnrelabel, alpha)
    rng = rng or RRng(seed)
    if initial_K is None:
        initial_K = np.stack([rng.sample_int(K, N) for _ in range(chains)])   # sample(1:K, N, replace=T)
    initial_K = np.ascontiguousarray(np.asarray(initial_K, dtype=np.int32).reshape(chains, N))
    return _run(_lib.SAMPLER_COLLAPSED, X, nsamples, K, alpha, beta, gamma, a, b, burnin, relabel, burnrelabel, debug,
                chains, seed, device, precision, init_z=initial_K, replay=replay, compact_z=compact_z, probes=probes,
                chain_offset=chain_offset)


def gibbs_dp(data, nsamples, alpha=None, a=1, b=1, beta=0.5, gamma=0.5, burnin=None, relabel=False,
             burnrelabel=50, maxK=30, debug=False, *, chains=1, seed=0, device=0, precision="fp64", replay=None,
             compact_z=False, probes=(), chain_offset=0):
    """Collapsed Gibbs sampler for the DP (CRP) infinite mixture (R/utils.R:23-30 -> collapsed_gibbs_dp.cpp:27)."""
    X = _as_X(data)
    burnin, burnrelabel, alpha = _defaults(nsamples, burnin, burnrelabel, alpha)
    return _run(_lib.SAMPLER_DP, X, nsamples, maxK, alpha, beta, gamma, a, b, burnin, relabel, burnrelabel, debug,
                chains, seed, device, precision, replay=replay, compact_z=compact_z, probes=probes,
                chain_offset=chain_offset)
