"""Python face of the CPU oracle (oracle/wepp_oracle.cpp) — TEST INFRASTRUCTURE ONLY.

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs, never by the product package wepp_b200/.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libwepp_oracle.so")
REF_LIB_PATH = os.path.join(_HERE, "_ref", "libwepp_ref.so")
VP = C.c_void_p
_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            import subprocess
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-pthread",
                                   os.path.join(_HERE, "wepp_oracle.cpp"), "-o", LIB_PATH])
        lib = C.CDLL(LIB_PATH)
        lib.oracle_cartesian_map.restype = C.c_int
        lib.oracle_read_scores.restype = C.c_int
        lib.oracle_stack_muts.restype = C.c_int64
        lib.oracle_rescore.restype = C.c_int
        _lib = lib
    return _lib


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def _p(a):
    return None if a is None else a.ctypes.data_as(VP)


def cartesian_map(arena, reads, mapped=None, n_threads: int = 1, epp_cap: int = 2048, want_node: bool = True,
                  range_trees: bool = False, range_reads=None):
    """Restated wepp_filter::cartesian_map.  Returns dict(max_parsimony, multiplicity, score,
    counts, epp_off, epp_nodes, seconds_map).  range_trees: score through the reference's range trees
    (arena.cpp:68-169; same results, less work per read), built from the windows of `range_reads` (default: `reads`).
    seconds_map = wall time of the read loop + merge alone (the reference's own timer boundary)."""
    lib = load()
    n, r = arena.n_nodes, reads.n_reads
    a = (_c(arena.parent, np.int32), _c(arena.mut_off, np.int64), _c(arena.mut_pos, np.int32),
         _c(arena.mut_ref, np.uint8), _c(arena.mut_nuc, np.uint8))
    rd = (_c(reads.start, np.int32), _c(reads.end, np.int32), _c(reads.degree, np.int32),
          _c(reads.rm_off, np.int64), _c(reads.rm_pos, np.int32), _c(reads.rm_nuc, np.uint8))
    m = None if mapped is None else _c(mapped, np.uint8)
    mp = np.zeros(r, np.int32)
    mu = np.zeros(r, np.int32)
    sc = np.zeros(n, np.float64) if want_node else None
    ct = np.zeros((n, 50), np.int32) if want_node else None
    cap = int(min(r * min(epp_cap, n) + 1, 1 << 30))
    eo = np.zeros(r + 1, np.int64)
    en = np.zeros(cap, np.int32)
    rr = reads if range_reads is None else range_reads
    rs, re_ = _c(rr.start, np.int32), _c(rr.end, np.int32)
    secs = C.c_double(0.0)
    rc = lib.oracle_cartesian_map_ex(C.c_int32(n), *[_p(x) for x in a], C.c_int32(arena.genome_size), C.c_int64(r),
                                     *[_p(x) for x in rd], _p(m), C.c_int32(n_threads), _p(mp), _p(mu), _p(sc), _p(ct),
                                     C.c_int32(epp_cap), C.c_int64(cap), _p(eo), _p(en), C.c_int32(1 if range_trees else 0),
                                     C.c_int64(rr.n_reads), _p(rs), _p(re_), C.byref(secs))
    if rc != 0:
        raise RuntimeError("oracle EPP buffer overflow")
    return {"max_parsimony": mp, "multiplicity": mu, "score": sc, "counts": ct, "epp_off": eo,
            "epp_nodes": en[: int(eo[-1])], "seconds_map": float(secs.value)}


def read_scores(arena, start: int, end: int, rm_pos, rm_nuc):
    """Parsimony of one read against every node (restated recursion)."""
    lib = load()
    a = (_c(arena.parent, np.int32), _c(arena.mut_off, np.int64), _c(arena.mut_pos, np.int32),
         _c(arena.mut_ref, np.uint8), _c(arena.mut_nuc, np.uint8))
    rp, rn = _c(rm_pos, np.int32), _c(rm_nuc, np.uint8)
    out = np.zeros(arena.n_nodes, np.int32)
    lib.oracle_read_scores(C.c_int32(arena.n_nodes), *[_p(x) for x in a], C.c_int32(start), C.c_int32(end),
                           C.c_int32(rp.shape[0]), _p(rp), _p(rn), _p(out))
    return out


def stack_muts(arena, nodes):
    lib = load()
    a = (_c(arena.parent, np.int32), _c(arena.mut_off, np.int64), _c(arena.mut_pos, np.int32),
         _c(arena.mut_ref, np.uint8), _c(arena.mut_nuc, np.uint8))
    sel = _c(nodes, np.int32)
    cap = 1 << 16
    while True:
        off = np.zeros(sel.shape[0] + 1, np.int64)
        pos = np.zeros(cap, np.int32)
        nuc = np.zeros(cap, np.uint8)
        t = lib.oracle_stack_muts(C.c_int32(arena.n_nodes), *[_p(x) for x in a], C.c_int32(sel.shape[0]), _p(sel),
                                  C.c_int64(cap), _p(off), _p(pos), _p(nuc))
        if t >= 0:
            return off, pos[:t], nuc[:t]
        cap *= 4


def rescore(arena, reads, cand_nodes, want_dist: bool = True):
    """Restated EPP-over-candidates (haplotype::mutation_distance + arena.cpp:614-625)."""
    lib = load()
    st_off, st_pos, st_nuc = stack_muts(arena, cand_nodes)
    r, c = reads.n_reads, len(cand_nodes)
    rd = (_c(reads.start, np.int32), _c(reads.end, np.int32), _c(reads.rm_off, np.int64),
          _c(reads.rm_pos, np.int32), _c(reads.rm_nuc, np.uint8))
    md = np.zeros(r, np.int32)
    dist = np.zeros((r, c), np.int32) if want_dist else None
    off = np.zeros(r + 1, np.int64)
    idx = np.zeros(max(r * c, 1), np.int32)
    lib.oracle_rescore(C.c_int32(c), _p(st_off), _p(_c(st_pos, np.int32)), _p(_c(st_nuc, np.uint8)), C.c_int64(r),
                       *[_p(x) for x in rd], _p(md), _p(dist), _p(off), _p(idx))
    return md, dist, off, idx[: int(off[-1])]
