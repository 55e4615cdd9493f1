"""CPU restatement of wepp_filter::filter (TEST INFRASTRUCTURE ONLY; pure Python over the C oracle).

Follows the reference's own shape, src/WEPP/initial_filter.cpp:
  find_correspondents :241-282   cached EPP list (binary search) or mutation_distance == max_parismony
  remove_read         :284-342   cached EPPs, else single_read_tree under the CURRENT mapped mask; score -= delta
  singular_step       :344-366
  clear_neighbors     :368-385   arena::highest_scoring_neighbors (arena.cpp:209-249)
  step                :387-453
  filter              :455-506   incl. the neighbour expansion
and arena.hpp:16-31 (score_comparator), haplotype.hpp:123-185 (mutation_distance, full_score).
Parity pinned against the reference's own object code in tests/test_ref_live.py (oracle/_ref).
"""
from __future__ import annotations

import functools
import math

import numpy as np

import oracle

SCORE_EPSILON = 1e-9            # config.hpp:15
MAX_CACHED_EPP_SIZE = 2048      # config.hpp:9
READ_DIST_FACTOR_THRESHOLD = 0.5 / 100
MAX_PEAK_PEAK_MUTATION = 2
FREYJA_PEAKS_LIMIT = 5000
TOP_N = 10
MAX_PEAKS = 300
MAX_NEIGHBORS_WEPP = 50
NUM_RANGE_BINS = 50


class _Haps:
    def __init__(self, arena):
        self.a = arena
        n = arena.n_nodes
        self.children = [[] for _ in range(n)]
        for v in range(1, n):
            self.children[int(arena.parent[v])].append(v)
        off, pos, nuc = oracle.stack_muts(arena, np.arange(n, dtype=np.int32))
        self.stack = [list(zip(pos[off[v]:off[v + 1]].tolist(), nuc[off[v]:off[v + 1]].tolist())) for v in range(n)]

    def distance(self, a: int, comp, lo: int = 0, hi: int = 2 ** 31 - 1) -> int:
        """haplotype::mutation_distance(comp, min_pos, max_pos), haplotype.hpp:123-173."""
        st = [m for m in self.stack[a] if lo <= m[0] <= hi]
        i = j = muts = 0
        while i < len(st) or j < len(comp):
            if i == len(st):
                muts += comp[j][1] != 15
                j += 1
            elif j == len(comp):
                muts += 1
                i += 1
            elif st[i][0] < comp[j][0]:
                muts += 1
                i += 1
            elif st[i][0] > comp[j][0]:
                muts += comp[j][1] != 15
                j += 1
            else:
                muts += (st[i][1] != comp[j][1]) and (comp[j][1] != 15)
                i += 1
                j += 1
        return muts

    def hap_distance(self, a: int, b: int) -> int:
        return self.distance(a, self.stack[b])


def filter_peaks(arena, reads, leaf_count, ids, n_threads: int = 2):
    """Returns (peaks, neighbours): sorted arena indices, as wepp_filter::filter returns them."""
    n, r = arena.n_nodes, reads.n_reads
    haps = _Haps(arena)
    o = oracle.cartesian_map(arena, reads, None, n_threads=n_threads, epp_cap=MAX_CACHED_EPP_SIZE)
    max_pars, mult = o["max_parsimony"], o["multiplicity"]
    cache = [o["epp_nodes"][o["epp_off"][i]:o["epp_off"][i + 1]] for i in range(r)]
    score = o["score"].astype(np.float64).copy()
    bins = np.minimum(np.asarray(reads.start) // (arena.genome_size // NUM_RANGE_BINS), NUM_RANGE_BINS - 1)
    true_counts = np.bincount(bins, weights=np.asarray(reads.degree, np.float64), minlength=NUM_RANGE_BINS)
    active = int((true_counts != 0).sum())
    with np.errstate(divide="ignore", invalid="ignore"):
        prop = o["counts"] / true_counts[None, :]
    divergence = (prop > READ_DIST_FACTOR_THRESHOLD).sum(axis=1) / active
    orig = score.copy()
    mapped = np.zeros(n, bool)
    read_muts = [list(zip(reads.rm_pos[reads.rm_off[i]:reads.rm_off[i + 1]].tolist(),
                          reads.rm_nuc[reads.rm_off[i]:reads.rm_off[i + 1]].tolist())) for i in range(r)]

    def full(v):
        return score[v] * math.sqrt(divergence[v])

    def cmp(l, rr):   # score_comparator: "l before rr" -> negative
        el, er = full(l), full(rr)
        if abs(el - er) > SCORE_EPSILON:
            return -1 if el > er else 1
        if leaf_count[l] != leaf_count[rr]:
            return -1 if leaf_count[l] > leaf_count[rr] else 1
        if ids[l] == ids[rr]:
            return 0
        return -1 if ids[l] > ids[rr] else 1

    key = functools.cmp_to_key(cmp)

    def neighbours(pivot, radius):
        curr = pivot
        while arena.parent[curr] >= 0 and haps.hap_distance(pivot, int(arena.parent[curr])) <= radius:
            curr = int(arena.parent[curr])
        out = []
        stack = [curr]
        while stack:
            v = stack.pop()
            if haps.hap_distance(pivot, v) > radius:
                continue
            if not mapped[v]:
                out.append(v)
            stack.extend(reversed(haps.children[v]))
        return out

    remaining = set(range(r))
    peaks = set()

    def singular_step(hap):
        corr = []
        for read in sorted(remaining):
            c = cache[read]
            k = int(np.searchsorted(c, hap))
            if k < len(c) and c[k] == hap:
                corr.append(read)
            elif len(c) == mult[read]:
                continue
            elif haps.distance(hap, read_muts[read], int(reads.start[read]), int(reads.end[read])) == max_pars[read]:
                corr.append(read)
        for read in corr:
            if len(cache[read]) == mult[read]:
                epps = cache[read]
            else:
                one = oracle.cartesian_map(arena, reads.slice(read, read + 1), mapped.astype(np.uint8), epp_cap=n,
                                           want_node=False)
                epps = one["epp_nodes"]
            if len(epps):
                delta = float(reads.degree[read]) / ((1 + int(max_pars[read])) * int(mult[read]))
                score[epps] -= delta
            remaining.discard(read)

    def step(current):
        if not current:
            return True, current
        min_score = full(current[0])
        if min_score < SCORE_EPSILON:
            return True, current
        consideration = []
        for v in current:
            if not (abs(full(v) - min_score) < SCORE_EPSILON and len(consideration) < TOP_N
                    and len(consideration) + len(peaks) < MAX_PEAKS):
                break
            if all(haps.hap_distance(old, v) > MAX_PEAK_PEAK_MUTATION for old in consideration):
                consideration.append(v)
                mapped[v] = True
        peaks.update(consideration)
        for pivot in consideration:
            for v in neighbours(pivot, MAX_PEAK_PEAK_MUTATION):
                mapped[v] = True
        for node in consideration:
            singular_step(node)
        current = [v for v in current if not (mapped[v] or score[v] <= SCORE_EPSILON)]
        current.sort(key=key)
        return len(peaks) >= MAX_PEAKS or not remaining or not current, current

    current = sorted(range(n), key=key)
    done = False
    while not done:
        done, current = step(current)

    nbrs = set()
    for k in range(5):
        score[:] = orig
        mapped[:] = False
        curr = set()
        for pivot in sorted(peaks):
            ordered = sorted(neighbours(pivot, MAX_PEAK_PEAK_MUTATION + k), key=key)
            i = 0
            for v in ordered:
                if v in peaks or v in curr:
                    continue
                mapped[v] = True
                curr.add(v)
                i += 1
                if i == MAX_NEIGHBORS_WEPP:
                    break
        if abs(FREYJA_PEAKS_LIMIT - (len(curr) + len(peaks))) < abs(FREYJA_PEAKS_LIMIT - (len(nbrs) + len(peaks))):
            nbrs = curr
    return np.array(sorted(peaks), np.int32), np.array(sorted(nbrs), np.int32)
