#!/usr/bin/env python
"""Decode the reference's R fixtures (gzip'd RDX2 / XDR v2 serialisation) without R.

Only run in the build container (needs /root/reference); the decoded arrays are committed
under tests/golden/ so nothing at test/bench time reads the reference tree.

Handles the SEXP types these fixtures use: SYMSXP, LISTSXP, CHARSXP, LGLSXP, INTSXP, REALSXP,
STRSXP, VECSXP, NILVALUE, REFSXP (+ attribute pairlists).  Matrices are column-major with
`dim`/`dimnames`; NA_real_ is a NaN with payload 1954, NA_integer_ is INT_MIN.
"""
import gzip
import struct
import sys

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

    def item(self):
        flags = self.i32()
        typ = flags & 0xFF
        has_attr = bool(flags & 0x200)
        has_tag = bool(flags & 0x400)
        if typ == 254:  # NILVALUE
            return None
        if typ == 255:  # REFSXP
            return self.refs[(flags >> 8) - 1]
        if typ == 1:  # SYMSXP
            name = self.item()
            self.refs.append(("sym", name))
            return ("sym", name)
        if typ == 2:  # LISTSXP (pairlist) -> python list of (tag, value)
            out = []
            while True:
                attr = self.item() if has_attr else None
                tag = self.item() if has_tag else None
                car = self.item()
                out.append((tag[1] if tag else None, car))
                flags = self.i32()
                typ = flags & 0xFF
                if typ == 254:
                    break
                if typ != 2:
                    raise ValueError("unexpected cdr type %d" % typ)
                has_attr = bool(flags & 0x200)
                has_tag = bool(flags & 0x400)
            return out
        if typ == 9:  # CHARSXP
            n = self.i32()
            if n == -1:
                return None
            s = self.b[self.p:self.p + n].decode("utf-8", "replace")
            self.p += n
            return s
        if typ in (10, 13):  # LGLSXP / INTSXP
            n = self.i32()
            v = np.frombuffer(self.b, dtype=">i4", count=n, offset=self.p).astype(np.int32)
            self.p += 4 * n
            obj = v
        elif typ == 14:  # REALSXP
            n = self.i32()
            v = np.frombuffer(self.b, dtype=">f8", count=n, offset=self.p).astype(np.float64)
            v[np.isnan(v)] = np.nan      # R's NA_real_ is a SIGNALLING NaN (payload 1954): canonicalise
            self.p += 8 * n
            obj = v
        elif typ == 16:  # STRSXP
            n = self.i32()
            obj = [self.item() for _ in range(n)]
        elif typ == 19:  # VECSXP
            n = self.i32()
            obj = [self.item() for _ in range(n)]
        else:
            raise ValueError("unsupported SEXP type %d at %d" % (typ, self.p))
        if has_attr:
            attrs = dict(self.item())
            obj = _apply_attrs(obj, attrs)
        return obj


def _apply_attrs(obj, attrs):
    if isinstance(obj, list) and "names" in attrs and all(isinstance(n, str) for n in attrs["names"]):
        if not all(isinstance(o, str) or o is None for o in obj) or len(obj) == 0:
            return dict(zip(attrs["names"], obj))
    if isinstance(obj, np.ndarray) and "dim" in attrs:
        dim = tuple(int(d) for d in attrs["dim"])
        arr = obj.reshape(dim, order="F")
        dn = attrs.get("dimnames")
        return {"__matrix__": arr, "dimnames": dn}
    return obj


def read_r(path):
    raw = gzip.open(path, "rb").read()
    if raw[:5] == b"RDX2\n":
        raw = raw[5:]
        top_is_pairlist = True
    else:
        top_is_pairlist = False
    assert raw[:2] == b"X\n", "only XDR serialisation supported"
    r = _Reader(raw)
    r.p = 2
    version, _, _ = r.i32(), r.i32(), r.i32()
    assert version == 2, version
    top = r.item()
    return dict(top) if top_is_pairlist else top


def mat(obj):
    """(array, rownames, colnames) from a decoded matrix."""
    a = obj["__matrix__"]
    dn = obj["dimnames"] or [None, None]
    return a, dn[0], dn[1]


def main(ref="/root/reference", out="tests/golden"):
    ex = read_r(f"{ref}/data/example_donor.RData")
    A, rn, cn = mat(ex["A_clone"])
    D, _, _ = mat(ex["D_clone"])
    Z, zr, zc = mat(ex["tree"]["Z"])
    P = ex["tree"]["P"]
    P = P["__matrix__"] if isinstance(P, dict) else P
    np.savez_compressed(
        f"{out}/example_donor.npz", A=A, D=D, Z=Z.astype(np.float64), P=np.asarray(P, float).ravel(),
        variants=np.array(rn), cells=np.array(cn), clones=np.array(zc), Z_variants=np.array(zr))
    print("example_donor", A.shape, Z.shape, "nnz", int(np.sum(np.nan_to_num(D) > 0)))

    si = read_r(f"{ref}/data/simulation_input.RData")
    Din, dr, dc = mat(si["D_input"])
    kw = {"D_input": Din}
    for name in ("tree_3clone", "tree_4clone", "tree_5clone"):
        if name in si:
            Zt, _, _ = mat(si[name]["Z"])
            kw[name + "_Z"] = Zt.astype(np.float64)
    np.savez_compressed(f"{out}/simulation_input.npz", **kw)
    print("simulation_input", Din.shape, {k: v.shape for k, v in kw.items()})

    # MatrixMarket AD / DP (dense-coverage de-novo case used by the vignette's mtDNA section)
    def read_mtx(path):
        with open(path) as f:
            lines = [ln for ln in f if not ln.startswith("%")]
        nr, nc, _ = (int(x) for x in lines[0].split())
        m = np.zeros((nr, nc))
        for ln in lines[1:]:
            i, j, v = ln.split()
            m[int(i) - 1, int(j) - 1] = float(v)
        return m
    ad = read_mtx(f"{ref}/inst/extdata/passed_ad.mtx")
    dp = read_mtx(f"{ref}/inst/extdata/passed_dp.mtx")
    names = [ln.strip() for ln in open(f"{ref}/inst/extdata/passed_variant_names.txt")]
    np.savez_compressed(f"{out}/mito_passed.npz", AD=ad, DP=dp, variants=np.array(names))
    print("mito", ad.shape, dp.shape, "max depth", dp.max())

    can = read_r(f"{ref}/inst/extdata/canopy_results.coveraged.rds")
    Zc, zr2, zc2 = mat(can["tree"]["Z"])
    np.savez_compressed(f"{out}/canopy_tree.npz", Z=Zc.astype(np.float64), variants=np.array(zr2), clones=np.array(zc2))
    print("canopy", Zc.shape)


if __name__ == "__main__":
    main(*sys.argv[1:])
