"""Minimal reader for R `.rda` files (bz2/gzip/xz + `RDX2\\nX\\n` XDR serialisation v2).

TEST INFRASTRUCTURE ONLY.  Used by `tests/golden/make_golden.py` (run in the build
container, where `/root/reference` is mounted) to turn the reference's bundled
datasets (`data/*.rda`, documented in reference `R/data.R:1-115`) into small
fixtures.  Nothing in the product (`gpedm_b200/`) imports this module.

Supported SEXP types: NILSXP(254), SYMSXP(1), LISTSXP(2), CHARSXP(9), LGLSXP(10),
INTSXP(13), REALSXP(14), STRSXP(16), VECSXP(19), REFSXP(255) - enough for data frames
of numeric / integer / factor / character columns.
"""
import bz2
import gzip
import lzma
import struct

import numpy as np


class _Reader:
    def __init__(self, buf):
        self.b = buf
        self.p = 0
        self.refs = []

    def i32(self):
        v = struct.unpack_from(">i", self.b, self.p)[0]
        self.p += 4
        return v

    def f64n(self, n):
        v = np.frombuffer(self.b, dtype=">f8", count=n, offset=self.p).astype(np.float64)
        self.p += 8 * n
        return v

    def i32n(self, n):
        v = np.frombuffer(self.b, dtype=">i4", count=n, offset=self.p).astype(np.int32)
        self.p += 4 * n
        return v

    def item(self):
        flags = self.i32()
        typ = flags & 0xFF
        has_attr = bool(flags & (1 << 9))
        has_tag = bool(flags & (1 << 10))
        if typ == 254:  # NILSXP
            return None
        if typ == 255:  # REFSXP
            idx = flags >> 8
            if idx == 0:
                idx = self.i32()
            return self.refs[idx - 1]
        if typ == 1:  # SYMSXP
            name = self.item()
            self.refs.append(name)
            return name
        if typ == 2:  # LISTSXP (pairlist): list of (tag, value)
            out = []
            while True:
                attr = self.item() if has_attr else None  # noqa: F841
                tag = self.item() if has_tag else None
                car = self.item()
                out.append((tag, car))
                nflags = self.i32()
                ntyp = nflags & 0xFF
                if ntyp == 254:
                    break
                if ntyp != 2:
                    raise ValueError("unsupported pairlist tail type %d" % ntyp)
                has_attr = bool(nflags & (1 << 9))
                has_tag = bool(nflags & (1 << 10))
            return out
        if typ == 9:  # CHARSXP
            n = self.i32()
            if n == -1:
                return None
            s = self.b[self.p:self.p + n].decode("utf-8", "replace")
            self.p += n
            return s
        if typ == 10 or typ == 13:  # LGLSXP / INTSXP
            n = self.i32()
            val = {"type": "int" if typ == 13 else "lgl", "data": self.i32n(n)}
        elif typ == 14:  # REALSXP
            n = self.i32()
            val = {"type": "real", "data": self.f64n(n)}
        elif typ == 16:  # STRSXP
            n = self.i32()
            val = {"type": "str", "data": [self.item() for _ in range(n)]}
        elif typ == 19:  # VECSXP
            n = self.i32()
            val = {"type": "list", "data": [self.item() for _ in range(n)]}
        else:
            raise ValueError("unsupported SEXP type %d at offset %d" % (typ, self.p))
        val["attr"] = dict(self.item()) if has_attr else {}
        return val


def _decompress(raw):
    if raw[:3] == b"BZh":
        return bz2.decompress(raw)
    if raw[:2] == b"\x1f\x8b":
        return gzip.decompress(raw)
    if raw[:6] == b"\xfd7zXZ\x00":
        return lzma.decompress(raw)
    return raw


def _column(v):
    """R vector -> numpy array (factor -> array of level strings, NA_integer_ -> nan)."""
    attr = v.get("attr", {})
    if v["type"] == "int" and "levels" in attr:
        lev = attr["levels"]["data"]
        return np.array([lev[i - 1] for i in v["data"]], dtype=object)
    if v["type"] == "int":
        d = v["data"].astype(np.float64)
        d[v["data"] == -2147483648] = np.nan
        return d
    if v["type"] == "real":
        return v["data"]
    if v["type"] == "str":
        return np.array(v["data"], dtype=object)
    raise ValueError("unsupported column type " + v["type"])


def read_rda(path):
    """Return {object_name: {column_name: numpy array}} for the data frames in `path`."""
    buf = _decompress(open(path, "rb").read())
    if buf[:7] != b"RDX2\nX\n":
        raise ValueError("not an RDX2 XDR file")
    r = _Reader(buf)
    r.p = 7
    r.i32(), r.i32(), r.i32()  # serialisation version, writer R version, min reader version
    top = r.item()  # pairlist of (name, object)
    out = {}
    for name, obj in top:
        names = obj["attr"]["names"]["data"]
        out[name] = {nm: _column(col) for nm, col in zip(names, obj["data"])}
    return out


if __name__ == "__main__":
    import sys
    for k, df in read_rda(sys.argv[1]).items():
        print(k, {c: (v.shape, v[:3]) for c, v in df.items()})
