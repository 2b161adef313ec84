#!/usr/bin/env python3
"""Decode the reference's bundled data/*.RData (gzip + RDX3 XDR) into plain-text fixtures.

Run in the build container only (needs /root/reference):
    python tools/decode_rdata.py
Writes tests/golden/<name>.txt: first line "N P", then N rows of P 0/1 integers.
Format notes: SURVEY.md Appendix C.  Reference docs: R/bmm-mcmc.R:10-55.
"""
import gzip, struct, sys, os

REF = "/root/reference/data"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


class Reader:
    def __init__(self, b):
        self.b, self.o = b, 0

    def i32(self):
        v = struct.unpack_from(">i", self.b, self.o)[0]
        self.o += 4
        return v

    def raw(self, n):
        v = self.b[self.o:self.o + n]
        self.o += n
        return v


def read_item(r):
    flags = r.i32()
    typ = flags & 0xFF
    has_attr = bool(flags & 0x200)
    has_tag = bool(flags & 0x400)
    if typ == 254:  # NILVALUE_SXP
        return None
    if typ == 2:  # LISTSXP (pairlist)
        out = []
        while True:
            attr = read_item(r) if has_attr else None
            tag = read_item(r) if has_tag else None
            car = read_item(r)
            out.append((tag, car))
            flags = r.i32()
            typ = flags & 0xFF
            has_attr = bool(flags & 0x200)
            has_tag = bool(flags & 0x400)
            if typ == 254:
                return out
            assert typ == 2, typ
    if typ == 1:  # SYMSXP
        return ("sym", read_item(r))
    if typ == 255:  # REFSXP
        return ("ref", flags >> 8)
    if typ == 9:  # CHARSXP
        n = r.i32()
        return r.raw(n).decode() if n >= 0 else None
    if typ == 13:  # INTSXP
        n = r.i32()
        vals = list(struct.unpack_from(">%di" % n, r.b, r.o))
        r.o += 4 * n
        attrs = read_item(r) if has_attr else None
        return ("int", vals, attrs)
    if typ == 16:  # STRSXP
        n = r.i32()
        return [read_item(r) for _ in range(n)]
    raise ValueError("unhandled SEXP type %d at %d" % (typ, r.o))


def decode(path):
    b = gzip.decompress(open(path, "rb").read())
    assert b[:5] == b"RDX3\n" and b[5:7] == b"X\n", b[:8]
    r = Reader(b)
    r.o = 7
    version, writer, minreader = r.i32(), r.i32(), r.i32()
    assert version == 3
    n = r.i32()
    r.raw(n)  # native encoding
    top = read_item(r)
    (tag, car), = top
    name = tag[1]
    kind, vals, attrs = car
    dim = None
    for t, v in attrs:
        if t[1] == "dim":
            dim = v[1]
    return name, dim, vals, writer


def main():
    os.makedirs(OUT, exist_ok=True)
    for nm in ("K2_N100_P5", "K2_N1000_P5", "K3_N1000_P5"):
        name, (N, P), vals, writer = decode(os.path.join(REF, nm + ".RData"))
        assert name == nm and len(vals) == N * P
        with open(os.path.join(OUT, nm + ".txt"), "w") as f:
            f.write("%d %d\n" % (N, P))
            for i in range(N):  # stored column-major
                f.write(" ".join(str(vals[i + N * d]) for d in range(P)) + "\n")
        print(nm, N, P, "writer=0x%08x" % writer, "colmeans",
              [round(sum(vals[N * d:N * (d + 1)]) / N, 3) for d in range(P)])


if __name__ == "__main__":
    main()
