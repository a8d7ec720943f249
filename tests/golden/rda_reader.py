"""Minimal reader for R's XDR serialisation (.rda files), used ONLY to extract
golden vectors from the reference's own fixtures (/root/reference/data/*.rda).

Test infrastructure: this file is run in the build container by make_golden.py;
nothing on the GPU box imports it.  Format notes: SURVEY.md Appendix D.
"""
import bz2
import gzip
import lzma
import struct


class RObj:
    """Generic R object: python value + attributes."""

    def __init__(self, kind, value=None, attr=None):
        self.kind, self.value, self.attr = kind, value, attr or {}

    def __repr__(self):
        return f"RObj({self.kind}, attr={list(self.attr)})"


def _decompress(raw):
    if raw[:2] == b"\x1f\x8b":
        return gzip.decompress(raw)
    if raw[:3] == b"BZh":
        return bz2.decompress(raw)
    if raw[:6] == b"\xfd7zXZ\x00":
        return lzma.decompress(raw)
    return raw


class _Reader:
    def __init__(self, buf):
        self.b, self.o, self.refs = buf, 0, []

    def int(self):
        v = struct.unpack_from(">i", self.b, self.o)[0]
        self.o += 4
        return v

    def ints(self, n):
        v = list(struct.unpack_from(f">{n}i", self.b, self.o))
        self.o += 4 * n
        return v

    def dbls(self, n):
        v = list(struct.unpack_from(f">{n}d", self.b, self.o))
        self.o += 8 * n
        return v

    def length(self):
        n = self.int()
        if n == -1:
            hi, lo = self.int(), self.int()
            n = (hi << 32) + lo
        return n

    def attrs(self):
        """Read an attribute pairlist into a dict."""
        pl = self.item()
        out = {}
        while isinstance(pl, RObj) and pl.kind == "pairlist":
            for tag, car in pl.value:
                out[tag] = car
            break
        return out

    def item(self):
        flags = self.int()
        t = flags & 0xFF
        has_attr = bool(flags & (1 << 9))
        has_tag = bool(flags & (1 << 10))
        if t == 254:
            return None
        if t in (242, 241, 253, 252, 251, 247):
            return RObj({242: "emptyenv", 241: "baseenv", 253: "globalenv", 252: "unbound",
                         251: "missing", 247: "basens"}[t])
        if t == 255:
            idx = flags >> 8
            if idx == 0:
                idx = self.int()
            return self.refs[idx - 1]
        if t in (249, 250, 248):
            self.int()
            n = self.int()
            names = [self.item() for _ in range(n)]
            ob = RObj("namespace", names)
            self.refs.append(ob)
            return ob
        if t == 1:
            name = self.item()
            self.refs.append(name)
            return name
        if t == 4:
            self.int()
            ob = RObj("env")
            self.refs.append(ob)
            enclos, frame, hashtab, attr = self.item(), self.item(), self.item(), self.item()
            ob.value = (frame, hashtab)
            return ob
        if t in (2, 6, 3, 5, 17):
            # pairlist-like, iterative on cdr
            items, attr0 = [], {}
            first = True
            while True:
                attr = self.attrs() if has_attr else {}
                if first:
                    attr0, first = attr, False
                tag = self.item() if has_tag else None
                car = self.item()
                items.append((tag, car))
                # cdr
                flags = self.int()
                t2 = flags & 0xFF
                if t2 == 254:
                    break
                if t2 != t:
                    # non-list cdr (rare); rewind and read as item
                    self.o -= 4
                    items.append(("__cdr__", self.item()))
                    break
                has_attr = bool(flags & (1 << 9))
                has_tag = bool(flags & (1 << 10))
            return RObj("pairlist" if t == 2 else f"lang{t}", items, attr0)
        if t in (7, 8):
            n = self.int()
            s = self.b[self.o:self.o + n].decode()
            self.o += n
            return RObj("builtin", s)
        if t == 9:
            n = self.int()
            if n == -1:
                return None
            s = self.b[self.o:self.o + n].decode("utf-8", "replace")
            self.o += n
            return s
        if t == 238:
            info, state, attr = self.item(), self.item(), self.item()
            cls = info.value[0][1]
            ob = self._altrep(cls, state)
            if isinstance(attr, RObj) and attr.kind == "pairlist" and isinstance(ob, RObj):
                for tag, car in attr.value:
                    ob.attr[tag] = car
            return ob
        if t in (10, 13):
            n = self.length()
            ob = RObj("int" if t == 13 else "lgl", self.ints(n))
        elif t == 14:
            n = self.length()
            ob = RObj("real", self.dbls(n))
        elif t == 15:
            n = self.length()
            ob = RObj("cplx", self.dbls(2 * n))
        elif t == 16:
            n = self.length()
            ob = RObj("str", [self.item() for _ in range(n)])
        elif t in (19, 20):
            n = self.length()
            ob = RObj("list", [self.item() for _ in range(n)])
        elif t == 24:
            n = self.length()
            ob = RObj("raw", self.b[self.o:self.o + n])
            self.o += n
        elif t == 25:
            ob = RObj("S4")
        elif t == 21:
            raise NotImplementedError("bytecode")
        else:
            raise NotImplementedError(f"SEXP type {t} at {self.o}")
        if has_attr:
            ob.attr = self.attrs()
        return ob

    def _altrep(self, cls, state):
        if cls == "compact_intseq":
            n, start, step = state.value
            return RObj("int", [int(start + k * step) for k in range(int(n))])
        if cls == "compact_realseq":
            n, start, step = state.value
            return RObj("real", [start + k * step for k in range(int(n))])
        if cls.startswith("wrap_"):
            return state.value[0][1] if state.kind == "pairlist" else state.value[0]
        if cls == "deferred_string":
            arg = state.value[0][1]
            return RObj("str", [str(v) for v in arg.value])
        raise NotImplementedError(f"ALTREP {cls}")


def read_rda(path):
    """Return {name: RObj} for every object saved in an .rda file."""
    with open(path, "rb") as f:
        buf = _decompress(f.read())
    assert buf[:5] in (b"RDX2\n", b"RDX3\n"), buf[:5]
    r = _Reader(buf)
    r.o = 5
    assert buf[r.o:r.o + 2] == b"X\n"
    r.o += 2
    version = r.int()
    r.int()
    r.int()
    if version == 3:
        n = r.int()
        r.o += n
    top = r.item()
    return {tag: car for tag, car in top.value}


def list_get(ob, name):
    """Named element of an R list (VECSXP with names attr)."""
    names = ob.attr["names"].value
    return ob.value[names.index(name)]
