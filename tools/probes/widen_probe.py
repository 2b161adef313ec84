import ctypes as C, time, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from bmm_mcmc_b200 import _lib, api
L = _lib.lib()
L.bmm_widen_u8_i32.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
n = 1 << 30
src = np.random.default_rng(0).integers(1, 4, n, dtype=np.uint8)
print("cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
try:
    print("cgroup cpu.max", open("/sys/fs/cgroup/cpu.max").read().strip())
except Exception as e:
    print("no cgroup info", e)
for name, dst in (("pageable", np.empty(n, dtype=np.int32)), ("pinned", api._empty((n,), np.int32, True))):
    dst[:] = 0
    for T in (1, 2, 4, 6, 8, 12, 16):
        t0 = time.perf_counter()
        L.bmm_widen_u8_i32(src.ctypes.data, dst.ctypes.data, n, T)
        dt = time.perf_counter() - t0
        print(name, "T=%2d  %.1f ms  %.1f GB/s out" % (T, dt * 1e3, 4 * n / dt / 1e9))
    assert (dst[:1000] == src[:1000]).all()
