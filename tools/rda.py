"""Minimal reader for R's RDX2 (XDR, big-endian) workspace files -- just enough for catnet's data/*.rda.

Used only by tools/make_fixtures.py (fixture generation in the build container, where /root/reference exists).
Handles NILVALUE, REFSXP, SYMSXP, LISTSXP, CHARSXP, LGLSXP, INTSXP, REALSXP, STRSXP, VECSXP, S4SXP and
attribute pairlists.  Returns plain Python: dict for pairlists, RObj for anything carrying attributes.
"""
import bz2
import gzip
import lzma
import struct


class RObj:
    def __init__(self, kind, value, attrs):
        self.kind, self.value, self.attrs = kind, value, attrs or {}

    def __repr__(self):
        return "RObj(%s, attrs=%s)" % (self.kind, list(self.attrs))


class _Reader:
    def __init__(self, buf):
        self.b, self.p, self.refs = buf, 0, []

    def i32(self):
        v = struct.unpack_from(">i", self.b, self.p)[0]
        self.p += 4
        return v

    def f64(self):
        v = struct.unpack_from(">d", self.b, self.p)[0]
        self.p += 8
        return v

    def item(self):
        flags = self.i32()
        t = flags & 0xFF
        has_attr, has_tag = bool(flags & 0x200), bool(flags & 0x400)
        if t == 254:                      # NILVALUE
            return None
        if t == 255:                      # REFSXP
            return self.refs[(flags >> 8) - 1]
        if t == 1:                        # SYMSXP
            name = self.item()
            self.refs.append(name)
            return name
        if t == 2:                        # LISTSXP (pairlist) -> dict / list of (tag, value)
            out = {}
            n = 0
            while True:
                attrs = self.item() if has_attr else None
                tag = self.item() if has_tag else ("_%d" % n)
                out[tag] = self.item()
                n += 1
                flags = self.i32()
                t = flags & 0xFF
                if t == 254:
                    return out
                assert t == 2, t
                has_attr, has_tag = bool(flags & 0x200), bool(flags & 0x400)
        if t == 9:                        # CHARSXP
            n = self.i32()
            if n < 0:
                return None
            s = self.b[self.p:self.p + n].decode("latin-1")
            self.p += n
            return s
        if t in (10, 13):                 # LGLSXP / INTSXP
            n = self.i32()
            v = list(struct.unpack_from(">%di" % n, self.b, self.p))
            self.p += 4 * n
        elif t == 14:                     # REALSXP
            n = self.i32()
            v = list(struct.unpack_from(">%dd" % n, self.b, self.p))
            self.p += 8 * n
        elif t in (16, 19):               # STRSXP / VECSXP
            n = self.i32()
            v = [self.item() for _ in range(n)]
        elif t == 25:                     # S4SXP
            v = None
        else:
            raise ValueError("unsupported SEXP type %d at %d" % (t, self.p))
        attrs = self.item() if has_attr else None
        if attrs or t == 25:
            return RObj(t, v, attrs)
        return v


def read_rda(path):
    raw = open(path, "rb").read()
    for opener in (gzip.decompress, bz2.decompress, lzma.decompress):
        try:
            raw = opener(raw)
            break
        except Exception:
            continue
    assert raw[:5] == b"RDX2\n", raw[:8]
    assert raw[5:7] == b"X\n"
    r = _Reader(raw)
    r.p = 7
    r.i32(), r.i32(), r.i32()
    return r.item()
