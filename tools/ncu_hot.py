#!/usr/bin/env python3
"""Hot SASS regions of a kernel from `ncu --page source --csv`:  python tools/ncu_hot.py rep.ncu-rep [top]
Prints instruction-count share by opcode and the most-sampled instructions with their stall reasons."""
import csv
import subprocess
import sys
from collections import Counter


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    body = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
    ci = {h: i for i, h in enumerate(hdr)}
    ex, smp = ci["Instructions Executed"], ci["# Samples"]
    tot = sum(int(r[ex]) for r in body)
    tots = sum(int(r[smp]) for r in body)
    ops = Counter()
    for r in body:
        op = r[ci["Source"]].split()
        op = op[1] if op and op[0].startswith("@") else (op[0] if op else "?")
        ops[op.split(".")[0]] += int(r[ex])
    print("total warp instructions %d, samples %d, SASS lines %d" % (tot, tots, len(body)))
    print("by opcode:", ", ".join("%s %.1f%%" % (k, 100.0 * v / tot) for k, v in ops.most_common(18)))
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    print("top sampled instructions:")
    for idx in sorted(range(len(body)), key=lambda i: -int(body[i][smp]))[:top]:
        r = body[idx]
        st = sorted(((int(r[ci[c]]), c[6:]) for c in stall_cols), reverse=True)[:3]
        print("  #%4d %5.1f%% smp  exec %8s  %-60s %s" % (idx, 100.0 * int(r[smp]) / max(tots, 1), r[ex], r[ci["Source"]].strip()[:60],
                                                   " ".join("%s=%d" % (n, v) for v, n in st if v)))


if __name__ == "__main__":
    main()
