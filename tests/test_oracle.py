"""CPU suite: the oracle restatement against the dense brute-force checker on tiny cases."""
import numpy as np
import pytest

import oracle
from oracle import brute
from tests import cases


@pytest.mark.parametrize("seed", range(12))
def test_oracle_matches_bruteforce(seed):
    arena, reads, mapped = cases.tiny_case(seed)
    for m in (None, mapped):
        o = oracle.cartesian_map(arena, reads, m, epp_cap=arena.n_nodes)
        b = brute.cartesian_map(arena, reads, m)
        assert np.array_equal(o["max_parsimony"], b["max_parsimony"])
        assert np.array_equal(o["multiplicity"], b["multiplicity"])
        assert np.array_equal(o["counts"], b["counts"])
        np.testing.assert_allclose(o["score"], b["score"], rtol=1e-12, atol=0)
        for r in range(reads.n_reads):
            assert np.array_equal(o["epp_nodes"][o["epp_off"][r]:o["epp_off"][r + 1]], b["epp"][r])


@pytest.mark.parametrize("seed", range(4))
def test_per_node_scores_match_bruteforce(seed):
    arena, reads, _ = cases.tiny_case(100 + seed)
    sc = brute.scores(arena, reads)
    for r in range(0, reads.n_reads, 7):
        a, b = int(reads.rm_off[r]), int(reads.rm_off[r + 1])
        got = oracle.read_scores(arena, int(reads.start[r]), int(reads.end[r]), reads.rm_pos[a:b], reads.rm_nuc[a:b])
        assert np.array_equal(got, sc[r])


def test_oracle_threads_agree():
    arena, reads = cases.small_case()
    o1 = oracle.cartesian_map(arena, reads, n_threads=1)
    o4 = oracle.cartesian_map(arena, reads, n_threads=4)
    for k in ("max_parsimony", "multiplicity", "counts", "epp_off", "epp_nodes"):
        assert np.array_equal(o1[k], o4[k])
    np.testing.assert_allclose(o1["score"], o4["score"], rtol=1e-12)


def test_epp_cap_semantics():
    """Reads above the cap are not cached (reference initial_filter.cpp:191-196)."""
    arena, reads, _ = cases.tiny_case(3)
    o = oracle.cartesian_map(arena, reads, epp_cap=3)
    ln = o["epp_off"][1:] - o["epp_off"][:-1]
    assert np.array_equal(ln, np.where(o["multiplicity"] <= 3, o["multiplicity"], 0))


@pytest.mark.parametrize("seed", range(4))
def test_mutation_distance_equals_tree_parsimony_on_consistent_tree(seed):
    """haplotype::mutation_distance (haplotype.hpp:123-177) equals the tree parsimony whenever the
    MAT's ref_nuc is the FASTA reference — find_correspondents relies on it (initial_filter.cpp:270-276)."""
    arena, reads, _ = cases.tiny_case(200 + seed)
    cand = np.arange(arena.n_nodes, dtype=np.int32)
    md, dist, off, idx = oracle.rescore(arena, reads, cand)
    sc = brute.scores(arena, reads)
    assert np.array_equal(dist, sc)
    assert np.array_equal(md, sc.min(axis=1))
    for r in range(reads.n_reads):
        assert np.array_equal(idx[off[r]:off[r + 1]], np.flatnonzero(sc[r] == sc[r].min()))
