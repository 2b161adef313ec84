"""Scratch: phase stamps of the tensor sweep kernel at a given N (rows on one GPU)."""
import ctypes as C, sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
import bench
from bmm_mcmc_b200 import _lib, api
N = int(sys.argv[1])
X = bench.synth_rows(0, N, 64, 32); ip, th = bench.grid_init(32, 64)
plan = api.Plan(_lib.SAMPLER_STICKBREAKING, X, 12, 32, chains=1, seed=1, init_pi=ip, init_theta=th, precision="fp32", compact_z=True,
                grid_path=True, alpha=1.0, beta=0.5, gamma=0.5, a=1.0, b=1.0, burnin=1, relabel=False, burnrelabel=0)
L = _lib.lib()
ns = 12
for r in range(3):
    plan.run(); plan.sync()
    t = (C.c_uint64 * 32)(); L.bmm_debug_ws_trace(t)
    t = [int(x) for x in t]
    pa, pb = (ns - 2) & 1, (ns - 1) & 1       # parity of the last two sweeps
    wa, wb, ua, ub = t[8*pa:8*pa+8], t[8*pb:8*pb+8], t[16+8*pa:16+8*pa+8], t[16+8*pb:16+8*pb+8]
    t0 = wa[0]
    f = lambda x: (x - t0) / 1e3
    print("   sweep B tail: flushed %.1f fence %.1f ticket %.1f exit %.1f" % (f(wb[4]), f(wb[6]), f(wb[7]), f(wb[5])))
    print("N=%d run %d | sweep A: entry 0.0 start %.1f first %.1f last %.1f flushed %.1f exit %.1f | upd A: b0 %.1f counts %.1f end %.1f / last blk %.1f end %.1f | sweep B: entry %.1f start %.1f first %.1f last %.1f flushed %.1f exit %.1f | upd B: b0 %.1f counts %.1f end %.1f / last %.1f end %.1f | kernel_ms/sweep %.4f" % (
        N, r, f(wa[1]), f(wa[2]), f(wa[3]), f(wa[4]), f(wa[5]), f(ua[0]), f(ua[1]), f(ua[2]), f(ua[3]), f(ua[4]),
        f(wb[0]), f(wb[1]), f(wb[2]), f(wb[3]), f(wb[4]), f(wb[5]), f(ub[0]), f(ub[1]), f(ub[2]), f(ub[3]), f(ub[4]), plan.kernel_ms()[0]/11))
c = (C.c_uint64 * 320)(); L.bmm_debug_ws_cta(c)
c = np.array([int(x) for x in c], dtype=np.uint64).reshape(160, 2)[:148]
ent = (c[:, 0] >> np.uint64(10)).astype(np.int64); sm = (c[:, 0] & np.uint64(1023)).astype(np.int64); end = (c[:, 1] & np.uint64(0x3FFFFFFFFFFFFF)).astype(np.int64)
t0 = ent.min()
dur = (end - ent) / 1e3
order = np.argsort(dur)
print("per-CTA duration (entry -> counts flushed) us: min %.1f median %.1f max %.1f; entry spread %.1f us" % (dur.min(), np.median(dur), dur.max(), (ent.max() - t0) / 1e3))
print("slowest 12: " + " ".join("cta%d/sm%d:%.1f" % (i, sm[i], dur[i]) for i in order[-12:]))
print("fastest 12: " + " ".join("cta%d/sm%d:%.1f" % (i, sm[i], dur[i]) for i in order[:12]))
print("by sm parity: even %.1f odd %.1f; sm<74 %.1f sm>=74 %.1f" % (dur[sm % 2 == 0].mean(), dur[sm % 2 == 1].mean(), dur[sm < 74].mean(), dur[sm >= 74].mean()))
plan.close()
