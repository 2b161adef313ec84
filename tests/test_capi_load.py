"""The C-ABI library: loads, exports every symbol include/bmm_capi.h declares, validates arguments
on the host, and refuses to compute without a CUDA device (no CPU fallback).  No GPU needed."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from bmm_mcmc_b200 import _lib
from conftest import ROOT, gpu_available


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "bmm_capi.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bmm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    syms = _header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), "libbmm_b200.so lacks " + s
    assert sorted(_lib.EXPORTS) == syms


def test_version_and_error_string():
    L = _lib.lib()
    assert b"sm_100a" in L.bmm_version()
    assert isinstance(L.bmm_last_error(), bytes)
    assert L.bmm_device_count() >= 0


def test_argument_validation_happens_before_cuda():
    L = _lib.lib()
    X = np.asfortranarray(np.zeros((10, 3), dtype=np.int32))
    a = _lib.Args()
    a.X = X.ctypes.data_as(C.POINTER(C.c_int32))
    a.N, a.P, a.nsamples, a.K = 10, 3, 1, 2   # nsamples too small
    a.beta = a.gamma = 0.5
    plan = C.c_void_p()
    rc = L.bmm_plan_create(_lib.SAMPLER_DP, C.byref(a), None, C.byref(plan))
    assert rc == -1 and b"nsamples" in L.bmm_last_error()
    a.nsamples = 10
    a.gamma = 0.7
    rc = L.bmm_plan_create(_lib.SAMPLER_DP, C.byref(a), None, C.byref(plan))
    assert rc == -4                               # collapsed_gibbs_dp.cpp:48-50
    a.gamma = 0.5
    a.relabel, a.burnin, a.burnrelabel = 1, 1, 1
    rc = L.bmm_plan_create(_lib.SAMPLER_DP, C.byref(a), None, C.byref(plan))
    assert rc == -1 and b"burnin" in L.bmm_last_error()


@pytest.mark.skipif(gpu_available(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback():
    import bmm_mcmc_b200 as B
    X = B.load_dataset("K2_N100_P5")
    with pytest.raises(_lib.BmmError) as e:
        B.gibbs_collapsed(X, 20, 2)
    assert e.value.code == -2


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "bmm_mcmc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in txt and "liboracle" not in txt and "oracle/" not in txt, f


@pytest.mark.parametrize("S,N,Cn,K,threads,nseg", [(187, 97, 5, 3, 1, 1), (1800, 40, 40, 3, 4, 3), (33, 1000, 64, 16, 3, 2),
                                                    (8, 8, 2, 1, 2, 1), (1800, 1000, 2, 3, 4, 4)])
def test_host_widening_derives_the_relabelled_matrix(S, N, Cn, K, threads, nseg):
    """Host half of the result download (host logic, no device): from the byte stream of z_original -- one buffer per
    sweep segment, laid out [chain][observation][sweeps of the segment] -- and the returned permutations, produce both
    int32 S x N column-major matrices, z being perm(s, z_original - 1) + 1 (full_gibbs.cpp:171-174); fed in chunks that
    split runs at arbitrary places.  Without permutations: plain widening of the segment."""
    L_ = _lib.lib()
    f = L_.bmm_widen_runs_u8_i32
    f.restype = None
    f.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                  C.c_void_p, C.c_int]
    rng = np.random.default_rng(S * 7 + N)
    zo = rng.integers(1, K + 1, size=(Cn, N, S), dtype=np.uint8)          # memory order of the S x N cm matrices
    perm = np.stack([np.stack([rng.permutation(K) for _ in range(S)], 0) for _ in range(Cn)], 0).astype(np.int32)  # [c][s][k]
    perm_cm = np.ascontiguousarray(perm.transpose(0, 2, 1))               # S x K column-major per chain
    want = perm[np.arange(Cn)[:, None, None], np.arange(S)[None, None, :], zo.astype(np.int64) - 1] + 1
    n = zo.size
    z = np.full(n + 8, -7, dtype=np.int32)
    o = np.full(n + 8, -7, dtype=np.int32)
    w = np.full(n + 8, -7, dtype=np.int32)
    bounds = [S * g // nseg for g in range(nseg + 1)]
    for g in range(nseg):
        s0, s1 = bounds[g], bounds[g + 1]
        seg = np.ascontiguousarray(zo[:, :, s0:s1]).reshape(-1)
        m = seg.size
        chunk = max(1, m // 3 + 5) if m < (1 << 20) else (1 << 20) + 64
        for lo in range(0, m, chunk):
            cnt = min(chunk, m - lo)
            piece = np.ascontiguousarray(seg[lo:lo + cnt])
            f(piece.ctypes.data, lo, cnt, s1 - s0, S, s0, N, K, perm_cm.ctypes.data, z.ctypes.data, o.ctypes.data, threads)
            f(piece.ctypes.data, lo, cnt, s1 - s0, S, s0, N, K, None, None, w.ctypes.data, threads)
    flat = zo.reshape(-1).astype(np.int32)
    assert np.array_equal(o[:n], flat) and np.array_equal(w[:n], flat)
    assert np.array_equal(z[:n], want.reshape(-1).astype(np.int32))
    assert (z[n:] == -7).all() and (o[n:] == -7).all() and (w[n:] == -7).all()
    # one output only
    z2 = np.zeros(n, dtype=np.int32)
    f(zo.ctypes.data, 0, n, S, S, 0, N, K, perm_cm.ctypes.data, z2.ctypes.data, None, threads)
    assert np.array_equal(z2, z[:n])
