"""Reader for R's ``save()`` files (RDX2 / RDX3, XDR encoding).

Test infrastructure: used once, in this container, by
``scripts/make_golden.py`` to turn the reference's bundled data
(``data/simdata.rda``, ``data/yeast_benchmark.rda``) and its test fixture
(``tests/testdata/ref_bces.rda``) into the ``.npz``/``.json`` files under
``tests/golden/``.  Nothing on the GPU box reads ``.rda`` files.

The format follows R's ``serialize.c`` (R 3.6): a header ``RDX{2,3}\\nX\\n``,
three version ints (+ native encoding string for version 3) and one
serialized object (a tagged pairlist name -> value).  Supported: all vector
types, attributes, S4 objects, environments / reference table, closures and
byte-code (``ref_bces.rda`` embeds NMF/Biclust objects carrying compiled
closures).
"""
from __future__ import annotations

import bz2
import gzip
import lzma
import struct

import numpy as np

NILVALUE, GLOBALENV, UNBOUND, MISSINGARG, BASENAMESPACE = 254, 253, 252, 251, 250
NAMESPACESXP, PACKAGESXP, PERSISTSXP = 249, 248, 247
EMPTYENV, BASEENV = 242, 241
ATTRLANGSXP, ATTRLISTSXP, ALTREP_SXP = 240, 239, 238
BCREPDEF, BCREPREF = 244, 243
REFSXP = 255
SYMSXP, LISTSXP, CLOSXP, ENVSXP, PROMSXP, LANGSXP = 1, 2, 3, 4, 5, 6
SPECIALSXP, BUILTINSXP, CHARSXP, LGLSXP, INTSXP, REALSXP = 7, 8, 9, 10, 13, 14
CPLXSXP, STRSXP, DOTSXP, VECSXP, EXPRSXP, BCODESXP = 15, 16, 17, 19, 20, 21
EXTPTRSXP, WEAKREFSXP, RAWSXP, S4SXP = 22, 23, 24, 25

NA_INT = -2147483648


class RObj:
    """A decoded R object: ``value`` plus ``attr`` dict (may be empty)."""

    __slots__ = ("kind", "value", "attr")

    def __init__(self, kind, value=None, attr=None):
        self.kind = kind
        self.value = value
        self.attr = attr or {}

    def __repr__(self):
        cls = self.attr.get("class")
        return f"<RObj {self.kind} class={cls.value if isinstance(cls, RObj) else cls}>"

    # convenience -----------------------------------------------------------
    def slot(self, name):
        return self.attr[name]

    def names(self):
        n = self.attr.get("names")
        return list(n.value) if n is not None else None

    def __getitem__(self, key):
        if isinstance(key, str):
            return self.value[self.names().index(key)]
        return self.value[key]

    def array(self):
        """Numeric vector/matrix as a NumPy array (column-major dims honoured)."""
        a = np.asarray(self.value)
        dim = self.attr.get("dim")
        if dim is not None:
            a = a.reshape(tuple(int(d) for d in dim.value), order="F")
        return a


class _Reader:
    def __init__(self, buf: bytes):
        self.b = buf
        self.p = 0
        self.refs = []

    def int(self):
        v = struct.unpack_from(">i", self.b, self.p)[0]
        self.p += 4
        return v

    def bytes(self, n):
        v = self.b[self.p:self.p + n]
        self.p += n
        return v

    def length(self):
        n = self.int()
        if n == -1:
            hi, lo = self.int(), self.int()
            n = (hi << 32) + (lo & 0xFFFFFFFF)
        return n

    def add_ref(self, obj):
        self.refs.append(obj)
        return obj

    # -- items --------------------------------------------------------------
    def item(self):
        flags = self.int()
        t = flags & 0xFF
        has_attr = bool(flags & (1 << 9))
        has_tag = bool(flags & (1 << 10))
        is_obj = bool(flags & (1 << 8))

        if t == NILVALUE:
            return None
        if t in (EMPTYENV, BASEENV, GLOBALENV, UNBOUND, MISSINGARG, BASENAMESPACE):
            return RObj({EMPTYENV: "emptyenv", BASEENV: "baseenv", GLOBALENV: "globalenv",
                         UNBOUND: "unbound", MISSINGARG: "missing",
                         BASENAMESPACE: "basenamespace"}[t])
        if t == REFSXP:
            idx = flags >> 8
            if idx == 0:
                idx = self.int()
            return self.refs[idx - 1]
        if t in (PERSISTSXP, PACKAGESXP, NAMESPACESXP):
            self.int()  # 0
            n = self.int()
            info = [self.item() for _ in range(n)]
            return self.add_ref(RObj({PERSISTSXP: "persist", PACKAGESXP: "package",
                                      NAMESPACESXP: "namespace"}[t], info))
        if t == SYMSXP:
            name = self.item()
            return self.add_ref(RObj("symbol", name.value if isinstance(name, RObj) else name))
        if t == ENVSXP:
            self.int()  # locked
            env = self.add_ref(RObj("env", {}))
            enclos = self.item()
            frame = self.item()
            hashtab = self.item()
            attr = self.item()
            d = {}
            if frame is not None:
                d.update(_pairlist_to_dict(frame))
            if hashtab is not None:
                for bucket in hashtab.value:
                    if bucket is not None:
                        d.update(_pairlist_to_dict(bucket))
            env.value = d
            if attr is not None:
                env.attr = _pairlist_to_dict(attr)
            return env
        if t in (LISTSXP, LANGSXP, CLOSXP, PROMSXP, DOTSXP):
            # iterate the cdr chain to avoid deep recursion on long pairlists
            head = None
            tail = None
            while True:
                attr = self.item() if has_attr else None
                tag = self.item() if has_tag else None
                car = self.item()
                node = [tag.value if isinstance(tag, RObj) and tag.kind == "symbol" else tag, car]
                cell = RObj({LISTSXP: "pairlist", LANGSXP: "lang", CLOSXP: "closure",
                             PROMSXP: "promise", DOTSXP: "dots"}[t], [node],
                            _pairlist_to_dict(attr) if attr is not None else None)
                if head is None:
                    head = cell
                    tail = cell
                else:
                    tail.value.append(node)
                # peek at the cdr: continue inline only for plain LISTSXP w/o attr
                save = self.p
                nflags = self.int()
                nt = nflags & 0xFF
                if nt == LISTSXP and t == LISTSXP and not (nflags & (1 << 9)):
                    has_tag = bool(nflags & (1 << 10))
                    has_attr = False
                    continue
                self.p = save
                cdr = self.item()
                if cdr is not None:
                    if isinstance(cdr, RObj) and cdr.kind in ("pairlist", "lang") and t != CLOSXP:
                        tail.value.extend(cdr.value)
                    else:
                        tail.value.append(["__cdr__", cdr])
                return head
        if t in (SPECIALSXP, BUILTINSXP):
            n = self.int()
            return RObj("builtin", self.bytes(n).decode("latin-1"))
        if t == CHARSXP:
            n = self.int()
            if n == -1:
                return RObj("char", None)
            return RObj("char", self.bytes(n).decode("latin-1"))
        if t == EXTPTRSXP:
            e = self.add_ref(RObj("extptr"))
            self.item()
            self.item()
            obj = e
        elif t == WEAKREFSXP:
            obj = self.add_ref(RObj("weakref"))
        elif t == LGLSXP:
            n = self.length()
            obj = RObj("logical", np.frombuffer(self.bytes(4 * n), dtype=">i4").astype(np.int32))
        elif t == INTSXP:
            n = self.length()
            obj = RObj("integer", np.frombuffer(self.bytes(4 * n), dtype=">i4").astype(np.int32))
        elif t == REALSXP:
            n = self.length()
            obj = RObj("double", np.frombuffer(self.bytes(8 * n), dtype=">f8").astype(np.float64))
        elif t == CPLXSXP:
            n = self.length()
            obj = RObj("complex", np.frombuffer(self.bytes(16 * n), dtype=">c16").astype(np.complex128))
        elif t == STRSXP:
            n = self.length()
            obj = RObj("character", [self.item().value for _ in range(n)])
        elif t in (VECSXP, EXPRSXP):
            n = self.length()
            obj = RObj("list" if t == VECSXP else "expression", [self.item() for _ in range(n)])
        elif t == RAWSXP:
            n = self.length()
            obj = RObj("raw", self.bytes(n))
        elif t == S4SXP:
            obj = RObj("S4")
        elif t == BCODESXP:
            nreps = self.int()
            reps = [None] * nreps
            obj = self.bc1(reps)
        elif t == ALTREP_SXP:
            info = self.item()
            state = self.item()
            attr = self.item()
            obj = RObj("altrep", (info, state), _pairlist_to_dict(attr) if attr is not None else None)
            return obj
        else:
            raise ValueError(f"unsupported SEXP type {t} at byte {self.p}")
        if has_attr:
            attr = self.item()
            obj.attr = _pairlist_to_dict(attr) if attr is not None else {}
        return obj

    # -- byte code ----------------------------------------------------------
    def bc1(self, reps):
        code = self.item()
        n = self.int()
        consts = []
        for _ in range(n):
            t = self.int()
            if t == BCODESXP:
                consts.append(self.bc1(reps))
            elif t in (LANGSXP, LISTSXP, BCREPDEF, BCREPREF, ATTRLANGSXP, ATTRLISTSXP):
                consts.append(self.bclang(t, reps))
            else:
                consts.append(self.item())
        return RObj("bytecode", (code, consts))

    def bclang(self, t, reps):
        if t == BCREPREF:
            return reps[self.int()]
        if t in (BCREPDEF, LANGSXP, LISTSXP, ATTRLANGSXP, ATTRLISTSXP):
            pos = -1
            if t == BCREPDEF:
                pos = self.int()
                t = self.int()
            has_attr = t in (ATTRLANGSXP, ATTRLISTSXP)
            node = RObj("lang" if t in (LANGSXP, ATTRLANGSXP) else "pairlist", [])
            if pos >= 0:
                reps[pos] = node
            if has_attr:
                a = self.item()
                node.attr = _pairlist_to_dict(a) if a is not None else {}
            tag = self.item()
            car = self.bclang(self.int(), reps)
            cdr = self.bclang(self.int(), reps)
            node.value = [[tag, car], ["__cdr__", cdr]]
            return node
        return self.item()


def _pairlist_to_dict(pl):
    out = {}
    if pl is None:
        return out
    for tag, car in pl.value:
        if tag == "__cdr__":
            continue
        out[tag] = car
    return out


def _decompress(raw: bytes) -> bytes:
    if raw[:2] == b"\x1f\x8b":
        return gzip.decompress(raw)
    if raw[:3] == b"BZh":
        return bz2.decompress(raw)
    if raw[:6] == b"\xfd7zXZ\x00":
        return lzma.decompress(raw)
    return raw


def load_rda(path: str) -> dict:
    """Return ``{name: RObj}`` for every object saved in ``path``."""
    with open(path, "rb") as fh:
        buf = _decompress(fh.read())
    magic = buf[:5]
    if magic not in (b"RDX2\n", b"RDX3\n"):
        raise ValueError(f"not an RDX2/RDX3 file: {magic!r}")
    r = _Reader(buf)
    r.p = 5
    if r.bytes(2) != b"X\n":
        raise ValueError("only XDR serialization is supported")
    version = r.int()
    r.int()
    r.int()
    if version == 3:
        n = r.int()
        r.bytes(n)
    top = r.item()
    return _pairlist_to_dict(top)
