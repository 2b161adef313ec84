"""Host widening throughput on this box: plain and relabel-deriving, AVX2 vs AVX-512 stores (BMM_WIDEN_ISA is read once per
process, so each ISA runs in its own subprocess), thread counts, run lengths."""
import ctypes as C, time, os, sys, subprocess
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if len(sys.argv) < 2:
    for rep in range(2):
        for isa in ("avx2", "avx512"):
            subprocess.run([sys.executable, __file__, isa], env=dict(os.environ, BMM_WIDEN_ISA=isa))
    sys.exit(0)
from bmm_mcmc_b200 import _lib, api
L = _lib.lib()
L.bmm_widen_u8_i32.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
f = L.bmm_widen_runs_u8_i32
f.restype = None
f.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
S, N, Cn, K = 1800, 1000, 256, 3
n = S * N * Cn
rng = np.random.default_rng(0)
src = rng.integers(1, K + 1, n, dtype=np.uint8)
perm = np.ascontiguousarray(np.stack([np.stack([rng.permutation(K) for _ in range(S)], 1) for _ in range(Cn)], 0).astype(np.int32))
try:
    z = api._empty((n,), np.int32, True); o = api._empty((n,), np.int32, True); kind = "pinned"
except Exception:
    z = np.empty(n, np.int32); o = np.empty(n, np.int32); kind = "pageable"
z[:] = 0; o[:] = 0
isa = sys.argv[1]
for T in (8, 12, 16):
    t0 = time.perf_counter(); L.bmm_widen_u8_i32(src.ctypes.data, o.ctypes.data, n, T); dt = time.perf_counter() - t0
    print("%s %s plain        T=%2d %.1f ms %.1f GB/s out" % (isa, kind, T, dt * 1e3, 4 * n / dt / 1e9))
    for Lr in (1800, 450):
        m = Cn * N * Lr
        t0 = time.perf_counter(); f(src.ctypes.data, 0, m, Lr, S, 0, N, K, perm.ctypes.data, z.ctypes.data, o.ctypes.data, T); dt = time.perf_counter() - t0
        print("%s %s derive L=%4d T=%2d %.1f ms %.1f GB/s out" % (isa, kind, Lr, T, dt * 1e3, 8 * m / dt / 1e9))
