"""ctypes binding of oracle/_ref/libbmm_ref.so = the reference's own UNMODIFIED C++ sources
(/root/reference/src/*.cpp) compiled against the header shim in oracle/shim/ (see oracle/build_ref.sh).

TEST INFRASTRUCTURE: import only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package (bmm_mcmc_b200) never imports this.

Calls go through the reference's registered `.Call` symbols (src/RcppExports.cpp:137-146) by name, with
the argument order of R/RcppExports.R:4-30, and come back as the R list the symbol returns
(dict of numpy arrays in R's column-major layout).  `set.seed` is `seed=`.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libbmm_ref.so")
_LIB = None

INTSXP, REALSXP, LGLSXP, VECSXP = 13, 14, 10, 19


class RefArg(C.Structure):
    _fields_ = [("type", C.c_int), ("ndim", C.c_int), ("dim", C.c_int * 3), ("data", C.c_void_p), ("n", C.c_longlong)]


def available():
    """True when the compiled reference is present (built here from /root/reference, shipped prebuilt to the GPU box)."""
    if os.path.exists(_SO):
        return True
    if os.path.isdir("/root/reference/src"):
        subprocess.check_call(["sh", os.path.join(_HERE, "build_ref.sh")], stdout=subprocess.DEVNULL)
    return os.path.exists(_SO)


def lib():
    global _LIB
    if _LIB is None:
        if not available():
            raise RuntimeError("oracle/_ref/libbmm_ref.so is not built and /root/reference is absent")
        L = C.CDLL(_SO)
        L.ref_last_error.restype = C.c_char_p
        L.ref_dotcall.restype = C.c_void_p
        L.ref_dotcall.argtypes = [C.c_char_p, C.c_int, C.POINTER(RefArg)]
        L.ref_free.argtypes = [C.c_void_p]
        for f in ("ref_type", "ref_ndim"):
            getattr(L, f).argtypes = [C.c_void_p]
        L.ref_length.argtypes = [C.c_void_p]
        L.ref_length.restype = C.c_longlong
        L.ref_dim.argtypes = [C.c_void_p, C.c_int]
        L.ref_data.argtypes = [C.c_void_p]
        L.ref_data.restype = C.c_void_p
        L.ref_list_name.argtypes = [C.c_void_p, C.c_int]
        L.ref_list_name.restype = C.c_char_p
        L.ref_list_elt.argtypes = [C.c_void_p, C.c_int]
        L.ref_list_elt.restype = C.c_void_p
        L.ref_registered_name.restype = C.c_char_p
        L.ref_registered_name.argtypes = [C.c_int, C.POINTER(C.c_int)]
        L.ref_update_alpha.restype = C.c_double
        L.ref_update_alpha.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int, C.c_int]
        _LIB = L
    return _LIB


def registered():
    """{symbol: arity} as registered by R_init_bmmmcmc (src/RcppExports.cpp:137-151)."""
    L = lib()
    out = {}
    for i in range(L.ref_n_registered()):
        n = C.c_int(0)
        name = L.ref_registered_name(i, C.byref(n)).decode()
        out[name] = n.value
    return out


def _arg(v, keep):
    a = RefArg()
    if isinstance(v, (bool, np.bool_)):
        arr = np.array([int(v)], dtype=np.int32)
        a.type = LGLSXP
    elif isinstance(v, (int, np.integer)):
        arr = np.array([v], dtype=np.int32)
        a.type = INTSXP
    elif isinstance(v, (float, np.floating)):
        arr = np.array([v], dtype=np.float64)
        a.type = REALSXP
    else:
        v = np.asarray(v)
        if v.dtype.kind in "iub":
            arr = np.asfortranarray(v, dtype=np.int32)
            a.type = INTSXP
        else:
            arr = np.asfortranarray(v, dtype=np.float64)
            a.type = REALSXP
        if arr.ndim >= 2:
            a.ndim = arr.ndim
            for i, d in enumerate(arr.shape):
                a.dim[i] = d
    keep.append(arr)
    a.data = arr.ctypes.data
    a.n = arr.size
    return a


def _unpack(L, h):
    t = L.ref_type(h)
    if t == VECSXP:
        out = {}
        for i in range(L.ref_length(h)):
            e = L.ref_list_elt(h, i)
            try:
                out[L.ref_list_name(h, i).decode()] = _unpack(L, e)
            finally:
                L.ref_free(e)
        return out
    n = L.ref_length(h)
    dt = np.float64 if t == REALSXP else np.int32
    buf = (C.c_char * (n * np.dtype(dt).itemsize)).from_address(L.ref_data(h)) if n else b""
    arr = np.frombuffer(buf, dtype=dt).copy()
    nd = L.ref_ndim(h)
    if nd:
        arr = arr.reshape([L.ref_dim(h, i) for i in range(nd)], order="F")
    return arr


def dotcall(symbol, *args, seed=None):
    """.Call(symbol, ...) on the compiled reference.  `seed` = set.seed(seed) first."""
    L = lib()
    if seed is not None:
        L.ref_set_seed(C.c_uint(seed))
    keep = []
    arr = (RefArg * len(args))(*[_arg(v, keep) for v in args])
    h = L.ref_dotcall(symbol.encode(), len(args), arr)
    if not h:
        raise RuntimeError("reference %s: %s" % (symbol, L.ref_last_error().decode()))
    try:
        return _unpack(L, h)
    finally:
        L.ref_free(h)


# ---- the generated R stubs (R/RcppExports.R:4-30), same names and argument order --------------------------
def gibbs_cpp(df, initialPi, initialTheta, nsamples, K, alpha, beta, gamma, a, b, burnin, relabel, burnrelabel,
              debug=False, seed=1):
    return dotcall("_bmmmcmc_gibbs_cpp", np.asarray(df, dtype=np.int32), np.asarray(initialPi, dtype=np.float64),
                   np.asarray(initialTheta, dtype=np.float64), int(nsamples), int(K), float(alpha), float(beta),
                   float(gamma), float(a), float(b), int(burnin), bool(relabel), int(burnrelabel), bool(debug), seed=seed)


def gibbs_stickbreaking_cpp(df, initialPi, initialTheta, nsamples, maxK, alpha, beta, gamma, a, b, burnin, relabel,
                            burnrelabel, debug=False, seed=1):
    return dotcall("_bmmmcmc_gibbs_stickbreaking_cpp", np.asarray(df, dtype=np.int32),
                   np.asarray(initialPi, dtype=np.float64), np.asarray(initialTheta, dtype=np.float64), int(nsamples),
                   int(maxK), float(alpha), float(beta), float(gamma), float(a), float(b), int(burnin), bool(relabel),
                   int(burnrelabel), bool(debug), seed=seed)


def collapsed_gibbs_cpp(df, initialK, nsamples, K, alpha, beta, gamma, a, b, burnin, relabel, burnrelabel,
                        debug=False, seed=1):
    return dotcall("_bmmmcmc_collapsed_gibbs_cpp", np.asarray(df, dtype=np.int32), np.asarray(initialK, dtype=np.int32),
                   int(nsamples), int(K), float(alpha), float(beta), float(gamma), float(a), float(b), int(burnin),
                   bool(relabel), int(burnrelabel), bool(debug), seed=seed)


def collapsed_gibbs_dp_cpp(df, nsamples, alpha, beta, gamma, a, b, burnin, relabel, burnrelabel, maxK, debug=False,
                           seed=1):
    return dotcall("_bmmmcmc_collapsed_gibbs_dp_cpp", np.asarray(df, dtype=np.int32), int(nsamples), float(alpha),
                   float(beta), float(gamma), float(a), float(b), int(burnin), bool(relabel), int(burnrelabel),
                   int(maxK), bool(debug), seed=seed)


def my_stephens_batch(p, debug=False):
    return dotcall("_bmmmcmc_my_stephens_batch", np.asarray(p, dtype=np.float64), bool(debug))


def my_lpsolve(x):
    return dotcall("_bmmmcmc_my_lpsolve", np.asarray(x, dtype=np.float64))


def rdirichlet_cpp(alpha_m, seed=1):
    return dotcall("_bmmmcmc_rdirichlet_cpp", np.asarray(alpha_m, dtype=np.float64).ravel(), seed=seed).ravel()


def my_stephens_online(q, p, sample_num):
    """src/stephens.cpp:66-94 (not registered with R; reached through src/stephens.h:6)."""
    L = lib()
    q = np.asfortranarray(q, dtype=np.float64)
    p = np.asfortranarray(p, dtype=np.float64)
    N, K = p.shape
    perm = np.zeros(K, dtype=np.int32)
    qn = np.zeros((N, K), order="F")
    rc = L.ref_stephens_online(N, K, q.ctypes.data_as(C.POINTER(C.c_double)), p.ctypes.data_as(C.POINTER(C.c_double)),
                               int(sample_num), perm.ctypes.data_as(C.POINTER(C.c_int)),
                               qn.ctypes.data_as(C.POINTER(C.c_double)))
    if rc:
        raise RuntimeError("reference my_stephens_online: %s" % L.ref_last_error().decode())
    return perm, qn


def update_alpha(alpha_old, a, b, N, K, seed=1):
    L = lib()
    L.ref_set_seed(C.c_uint(seed))
    return L.ref_update_alpha(alpha_old, a, b, N, K)
