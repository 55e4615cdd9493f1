"""Dense brute-force checker for tiny cases — TEST INFRASTRUCTURE ONLY.

Independent of the oracle's sparse merge: materialises, for every arena node, the allele
(and MAT ref_nuc) of the last mutation on its root path at every genome position, then counts
mismatches per SURVEY.md Appendix A (restating src/WEPP/initial_filter.cpp:64-66,118-123):

  no event on the path at p : mismatch iff the read carries a non-N mutation at p
  last event m at p         : a = read's mutation at p if any else m.ref_nuc;
                              mismatch iff a != N and a != m.mut_nuc
"""
from __future__ import annotations

import numpy as np

N = 15


def node_tables(arena):
    g, n = arena.genome_size, arena.n_nodes
    last_mut = np.zeros((n, g + 1), np.uint8)
    last_ref = np.zeros((n, g + 1), np.uint8)
    for v in range(n):
        p = int(arena.parent[v])
        if p >= 0:
            last_mut[v] = last_mut[p]
            last_ref[v] = last_ref[p]
        a, b = int(arena.mut_off[v]), int(arena.mut_off[v + 1])
        last_mut[v, arena.mut_pos[a:b]] = arena.mut_nuc[a:b]
        last_ref[v, arena.mut_pos[a:b]] = arena.mut_ref[a:b]
    return last_mut, last_ref


def read_codes(reads, r, g):
    c = np.zeros(g + 1, np.uint8)
    a, b = int(reads.rm_off[r]), int(reads.rm_off[r + 1])
    c[reads.rm_pos[a:b]] = reads.rm_nuc[a:b]
    return c


def scores(arena, reads, tables=None):
    """int32[R, N] parsimony of every read against every node."""
    last_mut, last_ref = tables if tables is not None else node_tables(arena)
    g = arena.genome_size
    out = np.zeros((reads.n_reads, arena.n_nodes), np.int32)
    for r in range(reads.n_reads):
        s, e = int(reads.start[r]), int(reads.end[r])
        if e < s:
            continue
        c = read_codes(reads, r, g)[s:e + 1][None, :]
        lm = last_mut[:, s:e + 1]
        lr = last_ref[:, s:e + 1]
        a = np.where(c == 0, lr, c)
        with_event = (a != N) & (a != lm)
        without = (c != 0) & (c != N)
        out[r] = np.where(lm != 0, with_event, without).sum(axis=1)
    return out


def cartesian_map(arena, reads, mapped=None):
    sc = scores(arena, reads)
    n = arena.n_nodes
    mapped = np.zeros(n, bool) if mapped is None else np.asarray(mapped).astype(bool)
    best = sc.min(axis=1)
    epp = [(np.flatnonzero((sc[r] == best[r]) & ~mapped)) for r in range(reads.n_reads)]
    mult = np.array([e.size for e in epp], np.int32)
    score = np.zeros(n, np.float64)
    counts = np.zeros((n, 50), np.int32)
    bin_size = arena.genome_size // 50
    for r in range(reads.n_reads):
        if mult[r] == 0:
            continue
        score[epp[r]] += float(reads.degree[r]) / ((1 + int(best[r])) * int(mult[r]))
        counts[epp[r], min(int(reads.start[r]) // bin_size, 49)] += int(reads.degree[r])
    return {"max_parsimony": best.astype(np.int32), "multiplicity": mult, "score": score, "counts": counts,
            "epp": epp, "scores": sc}
