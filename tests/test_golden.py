"""Golden fixtures produced by the reference's OWN object code (tests/golden/make_golden.py).

CPU part: pins the oracle restatement and the product's host-side arena builder against them.
GPU part (marked gpu): the CUDA path through the C ABI against the same fixtures."""
import os

import numpy as np
import pytest

import oracle
from tests.golden import make_golden as mg
from wepp_b200.placement import build_arena
from wepp_b200.synth import Arena, Reads

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = sorted(mg.CASES)


def load(name):
    g = np.load(os.path.join(HERE, name + ".npz"))
    tree, reads, masked, _ = mg.make_inputs(name)
    assert str(g["digest"]) == mg.digest(tree, reads, masked), "fixture inputs drifted: regenerate tests/golden"
    return g, tree, reads, masked


def golden_arena(g, tree, reads):
    arena = Arena(tree.genome_size, tree.ref_codes, g["arena_parent"], g["arena_mut_off"], g["arena_mut_pos"],
                  g["arena_mut_ref"], g["arena_mut_nuc"])
    mreads = Reads(reads.start, reads.end, reads.degree, g["reads_rm_off"], g["reads_rm_pos"], g["reads_rm_nuc"])
    return arena, mreads


@pytest.mark.parametrize("name", NAMES)
def test_arena_builder_matches_reference_arena(name):
    g, tree, reads, masked = load(name)
    arena, mreads, info = build_arena(tree, reads, masked)
    for k in ("parent", "source", "leaf_count", "mut_off", "mut_pos", "mut_ref", "mut_nuc"):
        assert np.array_equal(info[k], g["arena_" + k]), k
    assert np.array_equal(mreads.rm_off, g["reads_rm_off"])
    assert np.array_equal(mreads.rm_pos, g["reads_rm_pos"])
    assert np.array_equal(mreads.rm_nuc, g["reads_rm_nuc"])
    # folded-node CSR: every MAT node appears exactly once, each list starts with the source
    assert np.array_equal(np.sort(info["map_nodes"]), np.arange(tree.n_nodes))
    assert np.array_equal(info["map_nodes"][info["map_off"][:-1]], info["source"])


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_reference_cartesian_map(name):
    g, tree, reads, _ = load(name)
    arena, mreads = golden_arena(g, tree, reads)
    o = oracle.cartesian_map(arena, mreads, None, n_threads=4, epp_cap=2048)
    for k in ("max_parsimony", "multiplicity", "counts", "epp_off", "epp_nodes"):
        assert np.array_equal(o[k], g["cm_" + k]), k
    np.testing.assert_allclose(o["score"], g["cm_score"], rtol=1e-12, atol=0)


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_reference_masked_single_read_tree_and_distances(name):
    g, tree, reads, _ = load(name)
    arena, mreads = golden_arena(g, tree, reads)
    sel = g["srt_reads"]
    o = oracle.cartesian_map(arena, mreads.take(sel), g["srt_mapped"], epp_cap=arena.n_nodes, want_node=False)
    assert np.array_equal(o["max_parsimony"], g["srt_max_val"])
    assert np.array_equal(o["epp_off"], g["srt_off"]) and np.array_equal(o["epp_nodes"], g["srt_nodes"])
    md, dist, _, _ = oracle.rescore(arena, mreads, g["md_cand"])
    assert np.array_equal(dist, g["md_dist"])
    so, sp, sn = oracle.stack_muts(arena, np.arange(arena.n_nodes))
    assert np.array_equal(so, g["arena_st_off"]) and np.array_equal(sp, g["arena_st_pos"])
    assert np.array_equal(sn, g["arena_st_nuc"])


def hap_ids(source):
    """The reference's haplotype ids are its MAT node identifiers, "n<index>" in the fixture trees
    (oracle/ref_driver.cpp); the comparator's last tie-break compares them as strings (arena.hpp:29)."""
    ids = ["n%d" % int(s) for s in source]
    order = sorted(range(len(ids)), key=lambda i: ids[i])
    rank = np.zeros(len(ids), np.int32)
    rank[order] = np.arange(len(ids), dtype=np.int32)
    return ids, rank


@pytest.mark.parametrize("name", NAMES)
def test_restated_peak_loop_matches_reference_filter(name):
    """wepp_filter::filter (initial_filter.cpp:455-506) of the reference's object code vs oracle/peaks.py."""
    from oracle import peaks
    g, tree, reads, _ = load(name)
    arena, mreads = golden_arena(g, tree, reads)
    ids, _ = hap_ids(g["arena_source"])
    pk, nb = peaks.filter_peaks(arena, mreads, g["arena_leaf_count"], ids)
    assert pk.size > 0
    assert np.array_equal(np.sort(np.concatenate([pk, nb])), np.sort(g["filter_selected"]))


@pytest.mark.gpu
@pytest.mark.parametrize("env", [{}, {"WEPP_DELTA_PLACE": "2"}, {"WEPP_DELTA_PLACE": "2", "WEPP_PEAK_DELTA": "0"}],
                         ids=["default", "sparse_subsets", "sparse_map_list_subsets"])
@pytest.mark.parametrize("name", NAMES)
def test_cuda_peak_loop_matches_reference_filter(name, env, monkeypatch):
    """WEPP_DELTA_PLACE=2 forces the sparse corrections on these small read sets: the removed reads' per-node weights
    then come from the whole read set's plan (place_subset_by_delta) instead of a subset plan's window lists
    (WEPP_PEAK_DELTA=0 keeps those)."""
    from wepp_b200.placement import Placer
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    g, tree, reads, masked = load(name)
    arena, mreads, info = build_arena(tree, reads, masked)
    _, rank = hap_ids(info["source"])
    p = Placer(0)
    p.set_arena(arena)
    p.set_reads(mreads)
    pk, nb = p.filter_peaks(info["leaf_count"], rank)
    assert np.array_equal(np.sort(np.concatenate([pk, nb])), np.sort(g["filter_selected"]))
    # the cartesian_map state is restored afterwards (recover_haplotype_state)
    sc, ct = p.node_results()
    assert np.array_equal(ct, g["cm_counts"])
    np.testing.assert_allclose(sc, g["cm_score"], rtol=1e-9, atol=1e-15)
    p.close()


@pytest.mark.gpu
@pytest.mark.parametrize("force", [False, True], ids=["default", "sparse_subsets"])
@pytest.mark.parametrize("ranks", [2, 3])
@pytest.mark.parametrize("name", NAMES)
def test_cuda_group_peak_loop_matches_reference_filter(name, ranks, force, monkeypatch):
    """wepp_group_filter_peaks: the reads dealt over `ranks` ranks of one process (here all on device 0 — the ranks
    exchange through the same peer-memory kernel and event ordering as on separate GPUs; tests/test_multigpu_peer.py
    runs it on distinct devices): same peaks and neighbours as the reference's filter(), merged cartesian_map state
    identical on every rank."""
    from wepp_b200.multigpu import Group
    if force:
        monkeypatch.setenv("WEPP_DELTA_PLACE", "2")
    g, tree, reads, masked = load(name)
    arena, mreads, info = build_arena(tree, reads, masked)
    _, rank = hap_ids(info["source"])
    grp = Group([0] * ranks)
    grp.set_arena(arena)
    grp.set_reads(mreads)
    pk, nb = grp.filter_peaks(info["leaf_count"], rank)
    assert np.array_equal(np.sort(np.concatenate([pk, nb])), np.sort(g["filter_selected"]))
    sc0, ct0 = grp.node_results(0)
    assert np.array_equal(ct0, g["cm_counts"])
    np.testing.assert_allclose(sc0, g["cm_score"], rtol=1e-9, atol=1e-15)
    for r in range(1, ranks):
        sc, ct = grp.node_results(r)
        assert np.array_equal(ct, ct0) and np.array_equal(sc, sc0)   # bit-identical, not merely close
    grp.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_matches_reference_fixtures(name):
    from wepp_b200.placement import Placer, WeppFilter
    g, tree, reads, masked = load(name)
    arena, mreads, _ = build_arena(tree, reads, masked)
    p = Placer(0)
    p.set_arena(arena)
    f = WeppFilter(p).cartesian_map(mreads)
    assert np.array_equal(f.max_parismony, g["cm_max_parsimony"])
    assert np.array_equal(f.parsimony_multiplicity, g["cm_multiplicity"])
    assert np.array_equal(f.mapped_read_counts, g["cm_counts"])
    off, nodes = f.epp_positions_cache
    assert np.array_equal(off, g["cm_epp_off"]) and np.array_equal(nodes, g["cm_epp_nodes"])
    np.testing.assert_allclose(f.score, g["cm_score"], rtol=1e-9, atol=1e-15)
    np.testing.assert_allclose(f.dist_divergence, g["cm_dist_divergence"], rtol=1e-12, atol=0)
    # masked recompute of chosen reads (remove_read path) and candidate distances
    sel = g["srt_reads"].astype(np.int64)
    p.set_mapped(g["srt_mapped"])
    p.place_subset(sel, arena.n_nodes, int(g["srt_off"][-1]) + 16)
    mp, _ = p.read_results()
    off, nodes = p.epp()
    assert np.array_equal(mp[sel], g["srt_max_val"])
    for i, r in enumerate(sel):
        assert np.array_equal(nodes[off[r]:off[r + 1]], g["srt_nodes"][g["srt_off"][i]:g["srt_off"][i + 1]])
    md, dist, _, _ = p.rescore(g["md_cand"], want_dist=True)
    assert np.array_equal(dist, g["md_dist"])
    p.close()
