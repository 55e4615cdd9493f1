"""Host-side mirror of the reference's loaders over the C ABI (include/wepp_b200.h, "File formats"):
`load_mat` = dataset::mat() (src/WEPP/dataset.hpp:213-220), `load_reads` = load_reads_from_proto
(src/WEPP/sam2pb.cpp:489-549).  Parsing happens in libwepp_b200.so; this module only marshals."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import check, ptr


def _strings(off: np.ndarray, chars: np.ndarray) -> list[str]:
    b = chars.tobytes()
    return [b[off[i]:off[i + 1]].decode("utf-8", "replace") for i in range(off.shape[0] - 1)]


def pack_strings(strs) -> tuple[np.ndarray, np.ndarray]:
    enc = [s.encode() for s in strs]
    off = np.zeros(len(enc) + 1, np.int64)
    np.cumsum([len(e) for e in enc], out=off[1:])
    return off, np.frombuffer(b"".join(enc) or b"\0", dtype=np.uint8).copy()


@dataclass
class MatTree:
    """MAT::Tree as flat arrays, nodes in creation order (parent[v] < v, root 0)."""
    parent: np.ndarray
    mut_off: np.ndarray
    mut_pos: np.ndarray
    mut_ref: np.ndarray
    mut_par: np.ndarray
    mut_nuc: np.ndarray
    ids: list
    n_annotations: int
    clades: list            # per node: list of n_annotations strings
    genome_size: int = 0
    ref_codes: np.ndarray | None = None

    @property
    def n_nodes(self) -> int:
        return int(self.parent.shape[0])


def _mat_out(h) -> MatTree:
    lib = _lib.load()
    try:
        n, nm, na, ic, cc = C.c_int32(), C.c_int64(), C.c_int32(), C.c_int64(), C.c_int64()
        check(lib.wepp_mat_dims(h, C.byref(n), C.byref(nm), C.byref(na), C.byref(ic), C.byref(cc)))
        n, nm, na, ic, cc = n.value, nm.value, na.value, ic.value, cc.value
        parent = np.zeros(n, np.int32)
        mo = np.zeros(n + 1, np.int64)
        mp = np.zeros(nm, np.int32)
        mr, mpar, mn = (np.zeros(nm, np.uint8) for _ in range(3))
        io, ich = np.zeros(n + 1, np.int64), np.zeros(max(ic, 1), np.uint8)
        check(lib.wepp_mat_get(h, ptr(parent), ptr(mo), ptr(mp), ptr(mr), ptr(mpar), ptr(mn), ptr(io), ptr(ich)))
        co, cch = np.zeros(n * na + 1, np.int64), np.zeros(max(cc, 1), np.uint8)
        check(lib.wepp_mat_get_clades(h, ptr(co), ptr(cch)))
        flat = _strings(co, cch)
        clades = [flat[v * na:(v + 1) * na] for v in range(n)]
        return MatTree(parent, mo, mp, mr, mpar, mn, _strings(io, ich), na, clades)
    finally:
        lib.wepp_mat_free(h)


def load_mat(path: str, uncondense: bool = True) -> MatTree:
    h = C.c_void_p()
    check(_lib.load().wepp_mat_load(str(path).encode(), int(uncondense), C.byref(h)))
    return _mat_out(h)


def parse_mat(pb_bytes: bytes, uncondense: bool = True) -> MatTree:
    h = C.c_void_p()
    check(_lib.load().wepp_mat_parse(pb_bytes, len(pb_bytes), int(uncondense), C.byref(h)))
    return _mat_out(h)


def serialize_mat(parent, mut_off, mut_pos, mut_ref, mut_par, mut_nuc, ids) -> bytes:
    """Parsimony::data bytes (no condensed nodes, no metadata) of a flat tree."""
    lib = _lib.load()
    a = [np.ascontiguousarray(parent, np.int32), np.ascontiguousarray(mut_off, np.int64),
         np.ascontiguousarray(mut_pos, np.int32), np.ascontiguousarray(mut_ref, np.uint8),
         np.ascontiguousarray(mut_par, np.uint8), np.ascontiguousarray(mut_nuc, np.uint8)]
    io, ich = pack_strings(ids)
    n = check(lib.wepp_mat_serialize(a[0].shape[0], *[ptr(x) for x in a], ptr(io), ptr(ich), None, 0))
    buf = np.zeros(max(n, 1), np.uint8)
    check(lib.wepp_mat_serialize(a[0].shape[0], *[ptr(x) for x in a], ptr(io), ptr(ich), ptr(buf), n))
    return buf[:n].tobytes()


@dataclass
class ReadSet:
    """std::vector<raw_read> + dataset::read_reverse_merge."""
    names: list
    start: np.ndarray
    end: np.ndarray
    degree: np.ndarray
    rm_off: np.ndarray
    rm_pos: np.ndarray
    rm_nuc: np.ndarray
    reverse_merge: dict

    @property
    def n_reads(self) -> int:
        return int(self.start.shape[0])


def _reads_out(h) -> ReadSet:
    lib = _lib.load()
    try:
        d = [C.c_int64() for _ in range(7)]
        check(lib.wepp_reads_dims(h, *[C.byref(x) for x in d]))
        n, nm, nc, nk, nv, kc, vc = (x.value for x in d)
        start, end, degree = (np.zeros(n, np.int32) for _ in range(3))
        ro, rp, rn = np.zeros(n + 1, np.int64), np.zeros(nm, np.int32), np.zeros(nm, np.uint8)
        no, nch = np.zeros(n + 1, np.int64), np.zeros(max(nc, 1), np.uint8)
        check(lib.wepp_reads_get(h, ptr(start), ptr(end), ptr(degree), ptr(ro), ptr(rp), ptr(rn), ptr(no), ptr(nch)))
        ko, kch = np.zeros(nk + 1, np.int64), np.zeros(max(kc, 1), np.uint8)
        rev = np.zeros(nk + 1, np.int64)
        vo, vch = np.zeros(nv + 1, np.int64), np.zeros(max(vc, 1), np.uint8)
        check(lib.wepp_reads_get_reverse(h, ptr(ko), ptr(kch), ptr(rev), ptr(vo), ptr(vch)))
        keys, vals = _strings(ko, kch), _strings(vo, vch)
        rm = {k: vals[rev[i]:rev[i + 1]] for i, k in enumerate(keys)}
        return ReadSet(_strings(no, nch), start, end, degree, ro, rp, rn, rm)
    finally:
        lib.wepp_reads_free(h)


def load_reads(path: str, reference: str, n_threads: int = 4) -> ReadSet:
    h = C.c_void_p()
    ref = reference.encode()
    check(_lib.load().wepp_reads_load(str(path).encode(), ref, len(ref), int(n_threads), C.byref(h)))
    return _reads_out(h)


def parse_reads(pb_bytes: bytes, reference: str, n_threads: int = 4) -> ReadSet:
    h = C.c_void_p()
    ref = reference.encode()
    check(_lib.load().wepp_reads_parse(pb_bytes, len(pb_bytes), ref, len(ref), int(n_threads), C.byref(h)))
    return _reads_out(h)
