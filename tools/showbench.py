#!/usr/bin/env python3
"""Print the key numbers of bench JSON lines (scratch helper)."""
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "ERR", e); continue
    r = d.get("roofline", {})
    print("%s: value %.3e ms/step %.3f | e2e %.3e (%.1f ms) | launches %s | kernel_ms %.4f frac %.4f (%s) | n_gpus %s" % (
        f.split("/")[-1], d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"].get("ms_per_step", 0), d.get("gpu_launches"),
        r.get("kernel_ms", 0), r.get("frac", 0), r.get("bound"), d.get("n_gpus")))
    print("    extra.kernels_ms", d.get("extra", {}).get("kernels_ms"), "clk", d.get("clocks", {}).get("sm_mhz"), d.get("clocks", {}).get("samples"))
