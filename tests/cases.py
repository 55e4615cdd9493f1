"""Seeded small cases shared by the CPU and GPU suites (test infrastructure)."""
from __future__ import annotations

import numpy as np

from wepp_b200 import synth
from wepp_b200.synth import Arena, Reads


def tiny_case(seed: int, n_nodes: int = 40, genome: int = 64, n_reads: int = 60, max_len: int = 30):
    """Dense little tree + reads hitting every edge case the reference's code distinguishes:
    back-mutations (ref==mut), IUPAC mut_nuc, N in reads, read mutations where no node mutates,
    events exactly at window ends, all-N reads, empty-mutation reads, degree>1, root mutations."""
    rng = np.random.default_rng(seed)
    ref = np.zeros(genome + 1, np.uint8)
    ref[1:] = synth.ONE_HOT[rng.integers(0, 4, genome)]
    parent = np.full(n_nodes, -1, np.int32)
    for v in range(1, n_nodes):
        # preorder: parent on the rightmost path of v-1
        path = []
        u = v - 1
        while u >= 0:
            path.append(u)
            u = parent[u]
        parent[v] = path[int(rng.integers(0, len(path)))] if rng.random() < 0.6 else v - 1
    pos_l, ref_l, nuc_l, off = [], [], [], [0]
    for v in range(n_nodes):
        k = int(rng.integers(0 if v == 0 else 1, 4))
        ps = np.sort(rng.choice(np.arange(1, genome + 1), size=k, replace=False))
        for p in ps:
            r = rng.random()
            if r < 0.15:
                m = int(ref[p])                                   # back-mutation to ref
            elif r < 0.25:
                m = int(synth.IUPAC_AMBIG[rng.integers(0, 10)])    # ambiguity code
            else:
                m = int(synth.ONE_HOT[rng.integers(0, 4)])
            pos_l.append(int(p)); ref_l.append(int(ref[p])); nuc_l.append(m)
        off.append(len(pos_l))
    arena = Arena(genome, ref, parent, np.array(off, np.int64), np.array(pos_l, np.int32),
                  np.array(ref_l, np.uint8), np.array(nuc_l, np.uint8))
    st, en, dg, ro, rp, rn = [], [], [], [0], [], []
    for r in range(n_reads):
        ln = int(rng.integers(1, max_len + 1))
        s = int(rng.integers(1, genome - ln + 2))
        e = s + ln - 1
        kind = rng.random()
        muts = {}
        if kind < 0.1:
            muts = {p: 15 for p in range(s, e + 1)}              # all-N read
        elif kind < 0.2:
            muts = {}                                             # reference-identical read
        else:
            for p in range(s, e + 1):
                u = rng.random()
                if u < 0.12:
                    muts[p] = 15
                elif u < 0.30:
                    alts = [int(c) for c in synth.ONE_HOT if c != ref[p]]
                    muts[p] = alts[int(rng.integers(0, 3))]
        st.append(s); en.append(e); dg.append(int(rng.integers(1, 5)))
        for p in sorted(muts):
            rp.append(p); rn.append(muts[p])
        ro.append(len(rp))
    reads = Reads(np.array(st, np.int32), np.array(en, np.int32), np.array(dg, np.int32), np.array(ro, np.int64),
                  np.array(rp, np.int32), np.array(rn, np.uint8))
    mapped = (rng.random(n_nodes) < 0.2).astype(np.uint8)
    return arena, reads, mapped


def small_case(seed: int = 7, n_nodes: int = 3000, n_reads: int = 700, genome: int = 2000, **kw):
    arena = synth.make_arena(n_nodes, genome, seed, mean_depth=20.0, hot_sites=30, **kw)
    amps = synth.amplicon_scheme(genome, 12, 180, 260, seed)
    reads = synth.make_reads(arena, n_reads, seed, amplicons=amps, read_len=100, jitter=8, n_templates=50,
                             err=0.01, n_rate=0.02)
    return arena, reads


def star_case(seed: int = 0, n_leaves: int = 900, genome: int = 400, n_reads: int = 120):
    """A few internal nodes with hundreds of leaf children each, 1-3 events per leaf inside a
    small genome: long runs of point entries (some leaves with several events in one window)
    between boundary entries, so scan chunks must be cut at boundary entries."""
    rng = np.random.default_rng(seed)
    ref = np.zeros(genome + 1, np.uint8)
    ref[1:] = synth.ONE_HOT[rng.integers(0, 4, genome)]
    parent = [-1]
    hubs = [0]
    n_hubs = 4
    per = n_leaves // n_hubs
    for h in range(n_hubs):
        if h > 0:
            parent.append(hubs[int(rng.integers(0, len(hubs)))] if rng.random() < 0.5 else 0)
            # preorder needs the parent on the rightmost path: attach to the root or the previous hub chain
            parent[-1] = 0
            hubs.append(len(parent) - 1)
        hub = hubs[-1]
        for _ in range(per):
            parent.append(hub)
    parent = np.array(parent, np.int32)
    n_nodes = parent.shape[0]
    pos_l, ref_l, nuc_l, off = [], [], [], [0]
    for v in range(n_nodes):
        k = int(rng.integers(1, 4)) if v > 0 else 0
        ps = np.sort(rng.choice(np.arange(1, genome + 1), size=k, replace=False))
        for p in ps:
            alts = [int(c) for c in synth.ONE_HOT if c != ref[p]]
            pos_l.append(int(p)); ref_l.append(int(ref[p])); nuc_l.append(alts[int(rng.integers(0, 3))])
        off.append(len(pos_l))
    arena = Arena(genome, ref, parent, np.array(off, np.int64), np.array(pos_l, np.int32),
                  np.array(ref_l, np.uint8), np.array(nuc_l, np.uint8))
    amps = synth.amplicon_scheme(genome, 3, 150, 200, seed)
    reads = synth.make_reads(arena, n_reads, seed, amplicons=amps, read_len=120, jitter=6, n_templates=40,
                             err=0.01, n_rate=0.02)
    mapped = (rng.random(n_nodes) < 0.15).astype(np.uint8)
    return arena, reads, mapped
