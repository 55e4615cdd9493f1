"""Live pin: where oracle/_ref/libwepp_ref.so exists (built from /root/reference by `make -C oracle
ref`; it travels to the GPU box), the oracle and the arena builder are checked against the
reference's own object code on cases that are NOT in the committed fixtures."""
import numpy as np
import pytest

from oracle import ref

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (reference tree not mounted)")


def _live(seed):
    import oracle
    from tests import cases
    from wepp_b200.placement import build_arena
    tree, reads, _ = cases.tiny_case(seed) if seed < 1000 else (*cases.small_case(seed=seed, n_nodes=1500, n_reads=250), None)
    rng = np.random.default_rng(seed)
    masked = np.sort(rng.choice(np.arange(1, tree.genome_size + 1), 3, replace=False)).astype(np.int32)
    s = ref.Session(tree, reads, masked=masked, threads=2)
    a = s.arena()
    arena, mreads, info = build_arena(tree, reads, masked)
    ok = all(np.array_equal(a[k], info[k]) for k in ("parent", "source", "leaf_count", "mut_off", "mut_pos", "mut_nuc"))
    r = s.cartesian_map()
    o = oracle.cartesian_map(arena, mreads, None, n_threads=2, epp_cap=2048)
    ok = ok and all(np.array_equal(r[k], o[k]) for k in ("max_parsimony", "multiplicity", "counts", "epp_off", "epp_nodes"))
    ok = ok and np.allclose(r["score"], o["score"], rtol=1e-12, atol=0)
    return bool(ok)


@pytest.mark.parametrize("seed", [21, 22, 23, 1021])
def test_oracle_and_arena_builder_match_live_reference(seed):
    assert ref.run_case("tests.test_ref_live", "_live", seed)


def id_rank_from_source(source):
    """Rank of the reference's haplotype ids ("n<MAT node>", oracle/ref_driver.cpp) in ascending string order."""
    ids = ["n%d" % int(s) for s in source]
    order = sorted(range(len(ids)), key=lambda i: ids[i])
    rank = np.zeros(len(ids), np.int32)
    rank[order] = np.arange(len(ids), dtype=np.int32)
    return ids, rank


def _live_filter(seed, n_nodes, n_reads):
    """wepp_filter::filter of the reference's own object code vs the restated peak loop (oracle/peaks.py)."""
    from oracle import peaks
    from tests import cases
    from wepp_b200.placement import build_arena
    tree, reads = cases.small_case(seed=seed, n_nodes=n_nodes, n_reads=n_reads)
    s = ref.Session(tree, reads, masked=None, threads=2)
    got = np.sort(s.filter())
    arena, mreads, info = build_arena(tree, reads, None)
    ids, _ = id_rank_from_source(info["source"])
    pk, nb = peaks.filter_peaks(arena, mreads, info["leaf_count"], ids)
    want = np.sort(np.concatenate([pk, nb]))
    return got.tolist(), want.tolist(), int(pk.size)


@pytest.mark.parametrize("seed,n_nodes,n_reads", [(31, 600, 150), (32, 1200, 300), (33, 900, 400)])
def test_restated_peak_loop_matches_live_reference(seed, n_nodes, n_reads):
    got, want, n_peaks = ref.run_case("tests.test_ref_live", "_live_filter", seed, n_nodes, n_reads)
    assert n_peaks > 0
    assert got == want
