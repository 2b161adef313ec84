"""Builds tests/rhost/_build/librhost.so: r-package/src/host.cpp + driver.cpp against the Rcpp stand-in (oracle/shim/),
linked with libbmm_b200.so.  Test infrastructure; used by tests/test_rhost.py and __graft_entry__.build()."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SO = os.path.join(HERE, "_build", "librhost.so")


def build(force=False):
    src = [os.path.join(HERE, "driver.cpp"), os.path.join(ROOT, "r-package", "src", "host.cpp"),
           os.path.join(ROOT, "oracle", "shim", "RcppArmadillo.h"), os.path.join(ROOT, "include", "bmm_capi.h")]
    lib = os.path.join(ROOT, "bmm_mcmc_b200", "libbmm_b200.so")
    if not os.path.exists(lib):
        raise RuntimeError("libbmm_b200.so is not built")
    if force or not os.path.exists(SO) or any(os.path.getmtime(s) > os.path.getmtime(SO) for s in src + [lib]):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-I", os.path.join(ROOT, "oracle", "shim"),
                               "-I", os.path.join(ROOT, "oracle"), "-I", os.path.join(ROOT, "include"), src[0], "-o", SO,
                               "-L", os.path.join(ROOT, "bmm_mcmc_b200"), "-l:libbmm_b200.so",
                               "-Wl,-rpath," + os.path.join(ROOT, "bmm_mcmc_b200")])
    return SO
