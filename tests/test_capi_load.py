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
