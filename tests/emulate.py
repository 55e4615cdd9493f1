"""numpy emulation of the device algorithm on the HOST-built Euler stripes (test infrastructure).

Checks the product's host logic — signed-delta tables, boundary (enter/exit) and point (leaf)
entries, stripe grouping — against the oracle without a GPU: per read, gather the stripes its
window covers, order by sort key, prefix-sum the boundary deltas selected by the read's allele
code (the state every node inherits), add each leaf's point deltas to that leaf only, take the
min over nodes and count the unmapped nodes attaining it.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from wepp_b200 import _lib
from wepp_b200._lib import ptr


def host_stripes(arena, q: int):
    lib = _lib.load()
    a = [np.ascontiguousarray(x, dt) for x, dt in ((arena.parent, np.int32), (arena.mut_off, np.int64),
                                                   (arena.mut_pos, np.int32), (arena.mut_ref, np.uint8),
                                                   (arena.mut_nuc, np.uint8))]
    n_str = arena.genome_size // q + 1
    n = _lib.check(lib.wepp_host_euler_stripes(arena.n_nodes, *[ptr(x) for x in a], arena.genome_size, q, None, 0, None, 0))
    ent = np.zeros((max(n, 1), 4), np.uint32)
    off = np.zeros(n_str + 1, np.int64)
    _lib.check(lib.wepp_host_euler_stripes(arena.n_nodes, *[ptr(x) for x in a], arena.genome_size, q, ptr(ent), n,
                                           ptr(off), n_str + 1))
    return ent[:n], off


def delta_table(ent):
    """int8[n, 6]: delta for codes ref, A, C, G, T, N/outside."""
    z = ent[:, 2]
    d = np.zeros((ent.shape[0], 6), np.int8)
    for c in range(4):
        d[:, c] = ((z >> (8 * c)) & 0xFF).astype(np.uint8).view(np.int8)
    d[:, 4] = (ent[:, 3] & 0xFF).astype(np.uint8).view(np.int8)
    return d


CODE = {1: 1, 2: 2, 4: 3, 8: 4, 15: 5}


def emulate_place(arena, reads, mapped=None, q: int = 32):
    ent, off = host_stripes(arena, q)
    dt = delta_table(ent)
    n, g = arena.n_nodes, arena.genome_size
    mapped = np.zeros(n, bool) if mapped is None else np.asarray(mapped).astype(bool)
    mpre = np.concatenate([[0], np.cumsum(mapped)])
    best = np.zeros(reads.n_reads, np.int32)
    mult = np.zeros(reads.n_reads, np.int32)
    epps = []
    for r in range(reads.n_reads):
        s, e = int(reads.start[r]), int(reads.end[r])
        qs, qe = s // q, max(e, s) // q
        sl = slice(int(off[qs]), int(off[qe + 1]))
        en, d = ent[sl], dt[sl]
        order = np.argsort(en[:, 0], kind="stable")
        en, d = en[order], d[order]
        code = np.full(g + 2, 5, np.int64)
        if e >= s:
            code[s:e + 1] = 0
        a, b = int(reads.rm_off[r]), int(reads.rm_off[r + 1])
        for p, c in zip(reads.rm_pos[a:b], reads.rm_nuc[a:b]):
            code[p] = CODE[int(c)]
        seed = int(sum(1 for c in reads.rm_nuc[a:b] if c != 15))
        delta = d[np.arange(en.shape[0]), code[en[:, 1]]].astype(np.int64)
        idx = (en[:, 0] >> 1).astype(np.int64)
        point = (en[:, 0] & 1).astype(bool)
        # boundary entries: difference array over nodes; point entries: that leaf only
        diff = np.zeros(n + 1, np.int64)
        np.add.at(diff, idx[~point], delta[~point])
        score = seed + np.cumsum(diff[:n])
        np.add.at(score, idx[point], delta[point])
        m = score.min()
        nodes = np.flatnonzero((score == m) & ~mapped)
        best[r], mult[r] = m, nodes.size
        epps.append(nodes)
    return best, mult, epps
