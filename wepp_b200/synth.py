"""Seeded synthetic MAT arenas and collapsed read sets of the shapes BASELINE.json names.

Everything here produces *arena-level* inputs — the flattened, preorder-numbered condensed
tree and the `raw_read` fields (reference: src/WEPP/arena.cpp:3-56, src/WEPP/read.hpp:6-12)
— as plain numpy arrays, which is what both the C-ABI (include/wepp_b200.h) and the oracle
take.  The real quick-start data sets need a network download (reference README.md:231,256),
so the named configs are shape stand-ins and are labelled "synthetic" wherever reported.

Nucleotide codes follow the reference (src/mutation_annotated_tree.cpp:19-74): one-hot
A=1,C=2,G=4,T=8, IUPAC unions, N=15.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

SEED = 20260101
ONE_HOT = np.array([1, 2, 4, 8], dtype=np.uint8)
IUPAC_AMBIG = np.array([3, 5, 6, 9, 10, 12, 7, 11, 13, 14], dtype=np.uint8)
NUC_N = 15


@dataclass
class Arena:
    """Flattened condensed tree: node i is the i-th haplotype in preorder (arena index)."""
    genome_size: int
    ref_codes: np.ndarray      # uint8[G+1], one-hot reference base at 1-based position p (index 0 unused)
    parent: np.ndarray         # int32[N], parent[0] = -1, parent[v] < v
    mut_off: np.ndarray        # int64[N+1]
    mut_pos: np.ndarray        # int32[E], sorted and unique within a node
    mut_ref: np.ndarray        # uint8[E]  MAT ref_nuc
    mut_nuc: np.ndarray        # uint8[E]  MAT mut_nuc (may be an IUPAC union)

    @property
    def n_nodes(self) -> int:
        return int(self.parent.shape[0])

    @property
    def n_events(self) -> int:
        return int(self.mut_pos.shape[0])


@dataclass
class Reads:
    """Collapsed reads (`raw_read`): 1-based closed window, degree, sparse mutation list."""
    start: np.ndarray          # int32[R]
    end: np.ndarray            # int32[R]
    degree: np.ndarray         # int32[R]
    rm_off: np.ndarray         # int64[R+1]
    rm_pos: np.ndarray         # int32[M], sorted within a read, inside [start,end]
    rm_nuc: np.ndarray         # uint8[M], one of 1,2,4,8,15
    meta: dict = field(default_factory=dict)

    @property
    def n_reads(self) -> int:
        return int(self.start.shape[0])

    def slice(self, lo: int, hi: int) -> "Reads":
        a, b = int(self.rm_off[lo]), int(self.rm_off[hi])
        return Reads(self.start[lo:hi].copy(), self.end[lo:hi].copy(), self.degree[lo:hi].copy(),
                     (self.rm_off[lo:hi + 1] - a).copy(), self.rm_pos[a:b].copy(), self.rm_nuc[a:b].copy(),
                     dict(self.meta))

    def take(self, idx: np.ndarray) -> "Reads":
        idx = np.asarray(idx, dtype=np.int64)
        cnt = (self.rm_off[idx + 1] - self.rm_off[idx]).astype(np.int64)
        off = np.zeros(idx.size + 1, dtype=np.int64)
        np.cumsum(cnt, out=off[1:])
        src = np.repeat(self.rm_off[idx] - off[:-1], cnt) + np.arange(off[-1], dtype=np.int64)
        return Reads(self.start[idx].copy(), self.end[idx].copy(), self.degree[idx].copy(), off,
                     self.rm_pos[src].copy(), self.rm_nuc[src].copy(), dict(self.meta))


def make_reference(genome_size: int, rng: np.random.Generator) -> np.ndarray:
    ref = np.zeros(genome_size + 1, dtype=np.uint8)
    ref[1:] = ONE_HOT[rng.integers(0, 4, genome_size)]
    return ref


def make_arena(n_nodes: int, genome_size: int = 29903, seed: int = SEED, *, mean_depth: float = 80.0,
               extra_events: float = 0.25, hot_sites: int = 300, hot_frac: float = 0.2,
               iupac_frac: float = 0.001, revert_frac: float = 0.05, root_events: int = 0) -> Arena:
    """Random preorder tree: ~70 % leaves, depth reflected random walk with mean ~mean_depth;
    1+Poisson(extra_events) events on every non-root node (arena invariant: util.cpp:109);
    positions 80 % uniform / 20 % from a homoplasy hot list; ~5 % reversions to ref; rare IUPAC."""
    rng = np.random.default_rng(np.random.PCG64(seed))
    ref = make_reference(genome_size, rng)
    n = int(n_nodes)
    # depth process: +1 w.p. 0.3 (node v-1 becomes internal), else stay/pop; slight negative drift
    down = rng.random(n) < 0.3
    pops = rng.geometric(1.0 / (1.0 + 0.3 / 0.7 + 1.0 / (2.0 * mean_depth)), n) - 1
    step = np.where(down, 1, -pops).astype(np.int64)
    step[:2] = 0
    walk = np.cumsum(step)
    depth = walk - np.minimum.accumulate(np.minimum(walk, 0))  # reflect at 0
    if n > 1:
        depth[1:] = np.maximum(depth[1:], 0) + 1               # every non-root at depth >= 1
        # a node can be at most one level deeper than its predecessor
        # (guaranteed by step<=1 and reflection) — parent = last earlier node one level up
    depth[0] = 0
    parent = np.full(n, -1, dtype=np.int32)
    order = np.argsort(depth, kind="stable")
    dsorted = depth[order]
    bounds = np.searchsorted(dsorted, np.arange(dsorted[-1] + 2))
    for d in range(1, int(dsorted[-1]) + 1):
        up = order[bounds[d - 1]:bounds[d]]
        cur = order[bounds[d]:bounds[d + 1]]
        if cur.size == 0:
            continue
        k = np.searchsorted(up, cur) - 1
        parent[cur] = up[k]
    # events
    cnt = 1 + rng.poisson(extra_events, n)
    cnt[0] = root_events
    tot = int(cnt.sum())
    node_of = np.repeat(np.arange(n, dtype=np.int64), cnt)
    hot = rng.choice(np.arange(1, genome_size + 1), size=min(hot_sites, genome_size), replace=False)
    pos = rng.integers(1, genome_size + 1, tot)
    use_hot = rng.random(tot) < hot_frac
    pos[use_hot] = hot[rng.integers(0, hot.size, int(use_hot.sum()))]
    key = node_of * (genome_size + 2) + pos
    key = np.unique(key)                                        # sorted by (node, pos), duplicates dropped
    node_of = key // (genome_size + 2)
    pos = (key % (genome_size + 2)).astype(np.int32)
    # a non-root node that lost all events to de-duplication cannot happen (cnt>=1 keeps one)
    mut_off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(node_of, minlength=n), out=mut_off[1:])
    e = pos.size
    mref = ref[pos]
    shift = rng.integers(1, 4, e)
    ref_idx = np.log2(mref).astype(np.int64)
    mnuc = ONE_HOT[(ref_idx + shift) % 4]
    rev = rng.random(e) < revert_frac
    mnuc[rev] = mref[rev]
    amb = rng.random(e) < iupac_frac
    mnuc[amb] = IUPAC_AMBIG[rng.integers(0, IUPAC_AMBIG.size, int(amb.sum()))]
    return Arena(genome_size, ref, parent, mut_off, pos, mref.astype(np.uint8), mnuc.astype(np.uint8))


def haplotype_of(arena: Arena, node: int) -> tuple[np.ndarray, np.ndarray]:
    """Net root→node mutations (last event per position), as the reference's stack_muts
    (arena.cpp:18-46): positions sorted, alleles; entries equal to ref dropped."""
    path = []
    v = int(node)
    while v >= 0:
        path.append(v)
        v = int(arena.parent[v])
    last: dict[int, int] = {}
    for v in reversed(path):
        a, b = int(arena.mut_off[v]), int(arena.mut_off[v + 1])
        for k in range(a, b):
            if arena.mut_ref[k] != arena.mut_nuc[k]:
                last[int(arena.mut_pos[k])] = int(arena.mut_nuc[k])
            else:
                last.pop(int(arena.mut_pos[k]), None)
    ps = np.array(sorted(last), dtype=np.int32)
    return ps, np.array([last[p] for p in ps], dtype=np.uint8)


def amplicon_scheme(genome_size: int, n_amplicons: int, min_len: int, max_len: int, seed: int = SEED) -> np.ndarray:
    """Tiled, overlapping amplicon windows (1-based closed), ARTIC-like: int32[n,2]."""
    rng = np.random.default_rng(np.random.PCG64(seed + 17))
    lens = rng.integers(min_len, max_len + 1, n_amplicons)
    span = genome_size - 60 - int(lens[-1])
    starts = (30 + np.arange(n_amplicons) * (span / max(n_amplicons - 1, 1))).astype(np.int64)
    ends = np.minimum(starts + lens - 1, genome_size)
    return np.stack([starts, ends], axis=1).astype(np.int32)


def primer_scheme(name: str, min_len: int = 0, max_len: int = 1 << 30) -> np.ndarray:
    """Amplicon windows (1-based closed, int32[n,2]) of one of the reference's primer schemes — ARTICv4_1 (99 amplicons,
    388-493 bp), midnight (29, 1080-1223 bp), RSVA_all_primers_best_hits (50) — from wepp_b200/data/<name>.amplicons.tsv
    (derived from the reference's primers/<name>.bed by wepp_b200/data/make_amplicons.py).  Amplicons outside
    [min_len, max_len] are dropped (the RSV-A "best hits" file pairs two primers wrongly: a 96-bp and a 5.6-kb window)."""
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", name + ".amplicons.tsv")
    rows = [ln.split() for ln in open(path) if ln.strip() and not ln.startswith("#")]
    a = np.array([[int(r[1]), int(r[2])] for r in rows], dtype=np.int32)
    ln = a[:, 1] - a[:, 0] + 1
    return a[(ln >= min_len) & (ln <= max_len)]


def make_reads(arena: Arena, n_reads: int, seed: int = SEED, *, amplicons: np.ndarray | None = None,
               read_len: int = 150, jitter: int = 10, full_amplicon: bool = False, n_templates: int = 2000,
               err: float = 0.002, n_rate: float = 0.01, max_degree_geom: float = 0.5) -> Reads:
    """Reads drawn from `n_templates` leaf haplotypes.  Short-read mode: a `read_len` window at
    the left or right end of a random amplicon, start jittered by ±jitter.  full_amplicon=True:
    the read spans the whole amplicon (ONT-like).  Per-base substitution `err`, N rate `n_rate`."""
    rng = np.random.default_rng(np.random.PCG64(seed + 1))
    g = arena.genome_size
    n = arena.n_nodes
    if amplicons is None:
        amplicons = amplicon_scheme(g, 99, 388, 493, seed)
    # templates: leaves (a node is a leaf iff the next preorder node is not its child)
    is_leaf = np.ones(n, dtype=bool)
    is_leaf[arena.parent[1:]] = False
    leaves = np.flatnonzero(is_leaf)
    t_nodes = rng.choice(leaves, size=min(n_templates, leaves.size), replace=False)
    t_key_list, t_nuc_list = [], []
    for t, node in enumerate(t_nodes):
        ps, al = haplotype_of(arena, int(node))
        keep = (al & (al - 1)) == 0          # reads can only carry A/C/G/T
        t_key_list.append(t * (g + 2) + ps[keep].astype(np.int64))
        t_nuc_list.append(al[keep])
    t_keys = np.concatenate(t_key_list) if t_key_list else np.zeros(0, np.int64)
    t_nucs = np.concatenate(t_nuc_list) if t_nuc_list else np.zeros(0, np.uint8)

    r = int(n_reads)
    amp = rng.integers(0, amplicons.shape[0], r)
    a_s, a_e = amplicons[amp, 0].astype(np.int64), amplicons[amp, 1].astype(np.int64)
    if full_amplicon:
        start = a_s + rng.integers(0, jitter + 1, r)
        end = a_e - rng.integers(0, jitter + 1, r)
    else:
        left = rng.random(r) < 0.5
        jit = rng.integers(-jitter, jitter + 1, r)
        start = np.where(left, a_s + jit, a_e - read_len + 1 + jit)
        end = start + read_len - 1
    start = np.clip(start, 1, g)
    end = np.clip(end, start, g)
    length = end - start + 1
    tmpl = rng.integers(0, t_nodes.size, r)

    # source 1: template mutations inside the window
    lo = np.searchsorted(t_keys, tmpl * (g + 2) + start)
    hi = np.searchsorted(t_keys, tmpl * (g + 2) + end + 1)
    c1 = hi - lo
    rid1 = np.repeat(np.arange(r, dtype=np.int64), c1)
    src1 = np.repeat(lo - np.concatenate([[0], np.cumsum(c1)[:-1]]), c1) + np.arange(int(c1.sum()), dtype=np.int64)
    pos1 = (t_keys[src1] % (g + 2)).astype(np.int64)
    nuc1 = t_nucs[src1]
    # source 2: substitution errors; source 3: N
    def sprinkle(rate):
        k = rng.binomial(length, rate)
        rid = np.repeat(np.arange(r, dtype=np.int64), k)
        p = start[rid] + (rng.random(rid.size) * length[rid]).astype(np.int64)
        return rid, p
    rid2, pos2 = sprinkle(err)
    refi = np.log2(arena.ref_codes[pos2]).astype(np.int64)
    nuc2 = ONE_HOT[(refi + rng.integers(1, 4, pos2.size)) % 4]
    rid3, pos3 = sprinkle(n_rate)
    nuc3 = np.full(pos3.size, NUC_N, dtype=np.uint8)

    rid = np.concatenate([rid1, rid2, rid3])
    pos = np.concatenate([pos1, pos2, pos3])
    nuc = np.concatenate([nuc1, nuc2, nuc3]).astype(np.uint8)
    pri = np.concatenate([np.zeros(rid1.size, np.int8), np.ones(rid2.size, np.int8), np.full(rid3.size, 2, np.int8)])
    order = np.lexsort((-pri, pos, rid))                        # highest priority first within (read,pos)
    rid, pos, nuc = rid[order], pos[order], nuc[order]
    first = np.ones(rid.size, dtype=bool)
    first[1:] = (rid[1:] != rid[:-1]) | (pos[1:] != pos[:-1])
    rid, pos, nuc = rid[first], pos[first], nuc[first]
    # a substitution that lands on the reference base is not a mutation (cannot happen: shift>=1)
    rm_off = np.zeros(r + 1, dtype=np.int64)
    np.cumsum(np.bincount(rid, minlength=r), out=rm_off[1:])
    degree = rng.geometric(1.0 - max_degree_geom, r).astype(np.int32)
    return Reads(start.astype(np.int32), end.astype(np.int32), degree, rm_off, pos.astype(np.int32), nuc,
                 {"templates": t_nodes})


# ---- named shapes -------------------------------------------------------------
def config_shape(name: str, scale: float = 1.0, seed: int = SEED, n_reads: int | None = None,
                 read_seed: int | None = None) -> tuple[Arena, Reads, dict]:
    """C1..C4 stand-ins (BASELINE.md section 2 / SURVEY.md section 8d): synthetic trees, reads on the reference's own
    primer schemes.  `scale` shrinks node and read counts together (tests use tiny scales); n_reads overrides the
    read count (bench.py: one GPU's shard), read_seed the reads' seed (one per rank)."""
    rs = seed if read_seed is None else read_seed
    if name == "C1":      # RSV-A quick-start shape
        g, n, r = 15222, int(50_000 * scale), int(200_000 * scale)
        arena = make_arena(n, g, seed)
        reads = make_reads(arena, n_reads or r, rs, amplicons=primer_scheme("RSVA_all_primers_best_hits", 250, 600))
    elif name == "C2":    # SARS-CoV-2 quick-start shape
        g, n, r = 29903, int(1_000_000 * scale), int(1_000_000 * scale)
        arena = make_arena(n, g, seed)
        reads = make_reads(arena, n_reads or r, rs, amplicons=primer_scheme("ARTICv4_1"))
    elif name == "C3":    # public-scale MAT x ARTIC v4.1 reads
        g, n, r = 29903, int(8_000_000 * scale), int(10_000_000 * scale)
        arena = make_arena(n, g, seed)
        reads = make_reads(arena, n_reads or r, rs, amplicons=primer_scheme("ARTICv4_1"))
    elif name == "C4":    # ONT long reads, midnight primers
        g, n, r = 29903, int(8_000_000 * scale), int(1_000_000 * scale)
        arena = make_arena(n, g, seed)
        reads = make_reads(arena, n_reads or r, rs, amplicons=primer_scheme("midnight"),
                           full_amplicon=True, err=0.03, n_rate=0.05)
    else:
        raise ValueError(name)
    return arena, reads, {"name": name, "genome": g, "nodes": arena.n_nodes, "reads": reads.n_reads,
                          "events": arena.n_events, "seed": seed, "data": "synthetic"}
