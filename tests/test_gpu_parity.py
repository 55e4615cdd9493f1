"""GPU suite: the CUDA path through the C ABI against the oracle (bit-exact integers, scores to
1e-9 relative — the reference's own tie tolerance SCORE_EPSILON, src/WEPP/config.hpp:15)."""
import numpy as np
import pytest

import oracle
from tests import cases
from wepp_b200 import synth
from wepp_b200.placement import Placer, WeppFilter

pytestmark = pytest.mark.gpu

SCORE_RTOL = 1e-9   # doubles: relative tolerance (plus 1e-15 absolute for exact zeros)


def _check(arena, reads, mapped, q=None, k=None, epp_cap=None, threads=4):
    epp_cap = arena.n_nodes if epp_cap is None else epp_cap
    o = oracle.cartesian_map(arena, reads, mapped, n_threads=threads, epp_cap=epp_cap)
    p = Placer(0, stripe_width=q, reads_per_lane=k)
    p.set_arena(arena)
    p.set_reads(reads)
    p.set_mapped(mapped)
    p.place(epp_cap, int(o["epp_off"][-1]) + 16)
    mp, mu = p.read_results()
    sc, ct = p.node_results()
    off, nodes = p.epp()
    assert np.array_equal(mp, o["max_parsimony"])
    assert np.array_equal(mu, o["multiplicity"])
    assert np.array_equal(ct, o["counts"])
    np.testing.assert_allclose(sc, o["score"], rtol=SCORE_RTOL, atol=1e-15)
    assert np.array_equal(off, o["epp_off"])
    assert np.array_equal(nodes, o["epp_nodes"])
    st = p.stats()
    p.close()
    return st


@pytest.mark.parametrize("seed,q,k", [(0, 8, 8), (1, 8, 4), (2, 16, 2), (3, 4, 0), (4, 32, 8), (5, 1, 2)])
def test_tiny_cases_all_edge_kinds(seed, q, k):
    arena, reads, mapped = cases.tiny_case(seed)
    _check(arena, reads, None, q, k)
    _check(arena, reads, mapped, q, k)


@pytest.mark.parametrize("k", [0, 2, 4, 8])
def test_small_case(k):
    arena, reads = cases.small_case()
    st = _check(arena, reads, None, 32, k)
    assert st["n_tiles"] > 0 and st["kernel_launches"] >= 2


def test_small_case_with_mask_and_cap():
    arena, reads = cases.small_case(seed=11)
    rng = np.random.default_rng(5)
    mapped = (rng.random(arena.n_nodes) < 0.3).astype(np.uint8)
    _check(arena, reads, mapped, 32, 0, epp_cap=64)


@pytest.mark.parametrize("seed,k", [(0, 8), (1, 0), (2, 2)])
def test_star_tree_long_point_runs(seed, k):
    """Polytomies: hundreds of consecutive leaf (point) entries, multi-event leaves, mask, EPP lists."""
    arena, reads, mapped = cases.star_case(seed)
    _check(arena, reads, None, 32, k)
    _check(arena, reads, mapped, 16, k, epp_cap=200)


def test_empty_and_ragged_inputs():
    arena, reads = cases.small_case(seed=3, n_reads=70)
    # zero reads
    p = Placer(0)
    p.set_arena(arena)
    p.set_reads(reads.slice(0, 0))
    p.place(0, 0)
    sc, ct = p.node_results()
    assert not sc.any() and not ct.any()
    p.close()
    # one read, and a read count that is not a multiple of the tile size
    _check(arena, reads.slice(0, 1), None)
    _check(arena, reads.slice(0, 33), None, k=2)
    # single-node tree
    one = synth.Arena(arena.genome_size, arena.ref_codes, np.array([-1], np.int32), np.array([0, 0], np.int64),
                      np.zeros(0, np.int32), np.zeros(0, np.uint8), np.zeros(0, np.uint8))
    _check(one, reads, None)


def test_device_and_host_keying_agree_and_validate_reads():
    """wepp_set_reads keys the reads on the device (validation, bucket histogram, scatter); stripe
    geometries whose cell table would be too large are keyed on the host.  Both must give the
    oracle's results, and both must reject malformed reads with the same messages."""
    from wepp_b200._lib import WeppError
    arena, reads = cases.small_case(seed=7, n_reads=200)
    _check(arena, reads, None, 1, 0)      # stripe width 1 on a 2,000-base genome: 8M cells -> host keying
    _check(arena, reads, None, 32, 0)     # device keying
    for q in (1, 32):
        p = Placer(0, stripe_width=q)
        p.set_arena(arena)

        def bad(**kw):
            r = synth.Reads(reads.start.copy(), reads.end.copy(), reads.degree.copy(), reads.rm_off.copy(),
                            reads.rm_pos.copy(), reads.rm_nuc.copy())
            for name, (i, v) in kw.items():
                getattr(r, name)[i] = v
            return r
        has_mut = int(np.flatnonzero(np.diff(reads.rm_off) > 0)[0])
        k = int(reads.rm_off[has_mut])
        for r, msg in [(bad(start=(3, 0)), "read window"), (bad(end=(5, arena.genome_size + 1)), "read window"),
                       (bad(degree=(0, -1)), "degree"), (bad(rm_nuc=(k, 3)), "allele code"),
                       (bad(rm_pos=(k, int(reads.start[has_mut]) - 1)), "sorted, unique and inside")]:
            with pytest.raises(WeppError, match=msg):
                p.set_reads(r)
        p.set_reads(reads)                # the handle is usable again after a rejected set
        p.place(0, 0)
        mp, _ = p.read_results()
        assert np.array_equal(mp, oracle.cartesian_map(arena, reads, None, n_threads=4)["max_parsimony"])
        p.close()


def test_list_build_rank_table_and_binary_search_agree(monkeypatch):
    """The per-window Euler lists are built from the per-tree rank table when it fits, else by one binary
    search per stripe: both builds must place identically (and the table must survive a wider second read set)."""
    arena, reads = cases.small_case(seed=29)
    _check(arena, reads, None, 8, 0)
    monkeypatch.setenv("WEPP_NO_RANK_TABLE", "1")
    _check(arena, reads, None, 8, 0)
    monkeypatch.delenv("WEPP_NO_RANK_TABLE")
    p = Placer(0, stripe_width=8)
    p.set_arena(arena)
    short = reads.take(np.flatnonzero(reads.end - reads.start < np.median(reads.end - reads.start)))
    for r in (short, reads, short):      # the table is rebuilt when a read set spans more stripes
        p.set_reads(r)
        p.place(0, 0)
        mp, mu = p.read_results()
        o = oracle.cartesian_map(arena, r, None, n_threads=4)
        assert np.array_equal(mp, o["max_parsimony"]) and np.array_equal(mu, o["multiplicity"])
    p.close()


def test_scattered_windows_bucket_coarsening(monkeypatch):
    """Trimmed short reads (starts anywhere, lengths 30-150): merged buckets (reads served by a wider list of the
    same first stripe) and plain buckets must both give the oracle's placement, on the device- and host-keyed paths."""
    arena, base = cases.small_case(seed=17)
    rng = np.random.default_rng(6)
    idx = rng.integers(0, base.n_reads, 900)
    reads = base.take(idx)
    # shrink every window to a random sub-interval that still holds the read's mutations
    lo = reads.start.copy(); hi = reads.end.copy()
    for i in range(reads.n_reads):
        a, b = int(reads.rm_off[i]), int(reads.rm_off[i + 1])
        first = int(reads.rm_pos[a]) if b > a else int(hi[i])
        last = int(reads.rm_pos[b - 1]) if b > a else int(lo[i])
        lo[i] = rng.integers(lo[i], min(first, hi[i]) + 1)
        hi[i] = rng.integers(max(last, lo[i]), hi[i] + 1)
    reads = synth.Reads(lo.astype(np.int32), hi.astype(np.int32), reads.degree, reads.rm_off, reads.rm_pos, reads.rm_nuc)
    for q in (8, 1):      # 8: device keying, 1: host keying (cell table too large)
        st_m = _check(arena, reads, None, q, 0)
        monkeypatch.setenv("WEPP_NO_BUCKET_MERGE", "1")
        st_p = _check(arena, reads, None, q, 0)
        monkeypatch.delenv("WEPP_NO_BUCKET_MERGE")
        assert st_m["n_buckets"] < st_p["n_buckets"]


def test_everything_mapped_gives_zero_multiplicity():
    arena, reads = cases.small_case(seed=5, n_reads=100)
    mapped = np.ones(arena.n_nodes, np.uint8)
    o = oracle.cartesian_map(arena, reads, mapped)
    assert not o["multiplicity"].any()
    _check(arena, reads, mapped)


def test_medium_case_c2_shape():
    """SARS-CoV-2 shape at 1/50 scale: 20k nodes x 20k reads, genome 29,903."""
    arena, reads, _ = synth.config_shape("C2", scale=0.02)
    _check(arena, reads, None, epp_cap=2048, threads=8)


def test_long_reads_c4_shape():
    """ONT shape (1.1 kb windows) at small scale: exercises wide buckets / K=2."""
    arena = synth.make_arena(6000, 29903, 5)
    reads = synth.make_reads(arena, 300, 5, amplicons=synth.amplicon_scheme(29903, 29, 1058, 1201, 5),
                             full_amplicon=True, err=0.03, n_rate=0.05, n_templates=40)
    _check(arena, reads, None, epp_cap=2048, threads=8)


def test_place_subset_under_mask_matches_oracle():
    """remove_read's recompute path (initial_filter.cpp:298-302)."""
    arena, reads = cases.small_case(seed=13)
    rng = np.random.default_rng(1)
    mapped = (rng.random(arena.n_nodes) < 0.25).astype(np.uint8)
    sel = np.sort(rng.choice(reads.n_reads, 37, replace=False)).astype(np.int64)
    p = Placer(0)
    p.set_arena(arena)
    p.set_reads(reads)
    p.place(0, 0)
    base_mp, base_mu = p.read_results()
    p.set_mapped(mapped)
    p.place_subset(sel, arena.n_nodes, 37 * arena.n_nodes)
    mp, mu = p.read_results()
    off, nodes = p.epp()
    o = oracle.cartesian_map(arena, reads.take(sel), mapped, epp_cap=arena.n_nodes)
    assert np.array_equal(mp[sel], o["max_parsimony"]) and np.array_equal(mu[sel], o["multiplicity"])
    rest = np.setdiff1d(np.arange(reads.n_reads), sel)
    assert np.array_equal(mp[rest], base_mp[rest]) and np.array_equal(mu[rest], base_mu[rest])
    for i, r in enumerate(sel):
        assert np.array_equal(nodes[off[r]:off[r + 1]], o["epp_nodes"][o["epp_off"][i]:o["epp_off"][i + 1]])
    p.close()


def test_cartesian_map_host_call_and_filter_mirror():
    arena, reads = cases.small_case(seed=17)
    o = oracle.cartesian_map(arena, reads, None, n_threads=4)
    p = Placer(0)
    p.set_arena(arena)
    mp, mu, sc, ct = p.cartesian_map_host(reads)
    assert np.array_equal(mp, o["max_parsimony"]) and np.array_equal(mu, o["multiplicity"])
    assert np.array_equal(ct, o["counts"])
    np.testing.assert_allclose(sc, o["score"], rtol=SCORE_RTOL, atol=1e-15)
    f = WeppFilter(p).cartesian_map(reads)
    assert np.array_equal(f.max_parismony, mp)
    off, nodes = f.epp_positions_cache
    assert np.array_equal(off, o["epp_off"]) and np.array_equal(nodes, o["epp_nodes"])
    assert f.dist_divergence.shape == (arena.n_nodes,)
    # the device-side summary (wepp_get_node_summary) restates initial_filter.cpp:214-231 exactly
    sc2, dv = p.node_summary()
    assert np.array_equal(sc2, f.score)
    assert np.array_equal(dv, f.dist_divergence)
    assert 0.0 < dv.max() <= 1.0
    p.close()


def test_rescore_matches_oracle():
    arena, reads = cases.small_case(seed=19)
    rng = np.random.default_rng(2)
    cand = rng.choice(arena.n_nodes, 150, replace=False).astype(np.int32)
    p = Placer(0)
    p.set_arena(arena)
    p.set_reads(reads)
    md, dist, off, idx = p.rescore(cand, want_dist=True)
    omd, odist, ooff, oidx = oracle.rescore(arena, reads, cand)
    assert np.array_equal(md, omd) and np.array_equal(dist, odist)
    assert np.array_equal(off, ooff) and np.array_equal(idx, oidx)
    p.close()


def _rescore_case(name):
    if name.startswith("tiny"):
        arena, reads, _ = cases.tiny_case(int(name[4:]))
    elif name == "star":
        arena, reads, _ = cases.star_case(1)
    elif name == "c4":   # ONT shape: 1.1 kb windows, K = 2
        arena = synth.make_arena(6000, 29903, 5)
        reads = synth.make_reads(arena, 300, 5, amplicons=synth.amplicon_scheme(29903, 29, 1058, 1201, 5),
                                 full_amplicon=True, err=0.03, n_rate=0.05, n_templates=40)
    else:
        arena, reads = cases.small_case(seed=23)
    return arena, reads


@pytest.mark.parametrize("name,q,k,n_cand", [("tiny0", 8, 8, 5), ("tiny1", 8, 4, 40), ("tiny2", 16, 2, 3), ("tiny3", 4, 0, 1),
                                             ("tiny4", 32, 8, 64), ("star", 32, 8, 300), ("small", 32, 0, 1000),
                                             ("small", 16, 4, 9), ("c4", 32, 0, 257)])
def test_rescore_tile_kernel_matches_oracle(name, q, k, n_cand):
    """K4 over the resident reads (rescore_tiles.cuh): IUPAC / reversion / N edge cases of the tiny trees,
    leaves of a polytomy, long reads, candidate sets smaller than the warp count, repeated candidates."""
    arena, reads = _rescore_case(name)
    rng = np.random.default_rng(n_cand)
    cand = rng.integers(0, arena.n_nodes, n_cand).astype(np.int32)   # with repeats
    p = Placer(0, stripe_width=q, reads_per_lane=k)
    p.set_arena(arena)
    p.set_reads(reads)
    md, dist, off, idx = p.rescore(cand, want_dist=True)
    omd, odist, ooff, oidx = oracle.rescore(arena, reads, cand)
    assert np.array_equal(md, omd) and np.array_equal(dist, odist)
    assert np.array_equal(off, ooff) and np.array_equal(idx, oidx)
    md2, _, off2, idx2 = p.rescore(cand[::-1].copy())    # buffers are reused across calls
    omd2, _, ooff2, oidx2 = oracle.rescore(arena, reads, cand[::-1].copy())
    assert np.array_equal(md2, omd2) and np.array_equal(off2, ooff2) and np.array_equal(idx2, oidx2)
    md3, _, _, _ = p.rescore(cand, want_argmin=False)
    assert np.array_equal(md3, omd)
    md4, _, off4, _ = p.rescore(cand, want_argmin="count")   # compact entry lists: empty candidates evaluated in bulk
    assert np.array_equal(md4, omd) and np.array_equal(off4, ooff)
    p.close()


def _check_state_path(arena, reads, q=None, k=None, path=None):
    """place(0, 0) — no mask, no explicit EPP lists — against the oracle; `path` = the wepp_stats.place_path the
    call must have taken (0 Euler-list scan, 1 distinct states, 2 sparse corrections over the states)."""
    o = oracle.cartesian_map(arena, reads, None, n_threads=4, want_node=True)
    p = Placer(0, stripe_width=q, reads_per_lane=k)
    p.set_arena(arena)
    p.set_reads(reads)
    for _ in range(2):          # twice: the states are built once per read set and reused
        p.place(0, 0)
        mp, mu = p.read_results()
        sc, ct = p.node_results()
        assert np.array_equal(mp, o["max_parsimony"])
        assert np.array_equal(mu, o["multiplicity"])
        assert np.array_equal(ct, o["counts"])
        np.testing.assert_allclose(sc, o["score"], rtol=SCORE_RTOL, atol=1e-15)
        if path is not None:   # an int, or the set of paths allowed (tiny trees can overflow the states' position cap)
            assert p.stats()["place_path"] in (path if isinstance(path, tuple) else (path,))
    p.close()


@pytest.mark.parametrize("name,q,k", [("tiny0", 8, 8), ("tiny1", 8, 4), ("tiny2", 16, 2), ("tiny3", 4, 0), ("tiny4", 32, 8),
                                      ("tiny5", 1, 2), ("star", 32, 8), ("star", 16, 0), ("small", 32, 0), ("small", 16, 4),
                                      ("c4", 16, 0)])
def test_state_place_matches_oracle(name, q, k, monkeypatch):
    """state_place.cuh: scoring the distinct window-restricted haplotypes of every window list (dense over the states)."""
    monkeypatch.setenv("WEPP_STATE_PLACE", "1")
    monkeypatch.setenv("WEPP_DELTA_PLACE", "0")
    arena, reads = _rescore_case(name)
    _check_state_path(arena, reads, q, k, path=(0, 1) if name.startswith("tiny") or name == "c4" else 1)   # tiny / wide windows may exceed the states' position cap


@pytest.mark.parametrize("cand", [None, 0, 3])
@pytest.mark.parametrize("name,q,k", [("tiny0", 8, 8), ("tiny1", 8, 4), ("tiny2", 16, 2), ("tiny3", 4, 0), ("tiny4", 32, 8),
                                      ("tiny5", 1, 2), ("star", 32, 8), ("star", 16, 0), ("small", 32, 0), ("small", 16, 4)])
def test_delta_place_matches_oracle(name, q, k, cand, monkeypatch):
    """delta_place.cuh: sparse corrections per read over the states.  WEPP_DELTA_PLACE=2 forces the path on read sets
    with a window per read (the tiny cases: all-N reads and reads with dozens of mutations take the byte-scratch
    route); WEPP_DELTA_CAND shrinks the candidate queue so that the posting re-walk runs too."""
    monkeypatch.setenv("WEPP_DELTA_PLACE", "2")
    if (q + k) % 3 == 0:
        monkeypatch.setenv("WEPP_DELTA_CTAS", "1")   # one CTA of 16 warps per SM (the shape of lists too wide for half an SM) on a third of the cases
    if cand is not None:
        monkeypatch.setenv("WEPP_DELTA_CAND", str(cand))
    arena, reads = _rescore_case(name)
    _check_state_path(arena, reads, q, k, path=(0, 2) if name.startswith("tiny") else 2)


def test_delta_place_is_the_default_for_amplicon_reads():
    arena, reads = cases.small_case(seed=23, n_reads=4000)
    _check_state_path(arena, reads, path=2)


@pytest.mark.parametrize("name,q,k", [("tiny0", 8, 8), ("star", 32, 0), ("small", 16, 0)])
def test_euler_scan_accumulate_without_epp_lists(name, q, k, monkeypatch):
    """WEPP_STATE_PLACE=0: the same call through place_kernel's accumulate-only variant."""
    monkeypatch.setenv("WEPP_STATE_PLACE", "0")
    arena, reads = _rescore_case(name)
    _check_state_path(arena, reads, q, k)


def test_state_place_across_read_sets():
    """The states are kept while consecutive read sets map to the same window lists and buckets, and rebuilt
    otherwise: alternate between a read set, a subset of it and the set again."""
    arena, reads = cases.small_case(seed=31)
    sub = reads.take(np.arange(0, reads.n_reads, 3))
    p = Placer(0)
    p.set_arena(arena)
    for r in (reads, sub, reads, reads, sub):
        o = oracle.cartesian_map(arena, r, None, n_threads=4)
        p.set_reads(r)
        p.place(0, 0)
        mp, mu = p.read_results()
        sc, ct = p.node_results()
        assert np.array_equal(mp, o["max_parsimony"]) and np.array_equal(mu, o["multiplicity"])
        assert np.array_equal(ct, o["counts"])
        np.testing.assert_allclose(sc, o["score"], rtol=SCORE_RTOL, atol=1e-15)
    p.close()


@pytest.mark.parametrize("delta,path", [("0", 1), ("1", 2)])
def test_state_place_medium_c2_shape(delta, path, monkeypatch):
    monkeypatch.setenv("WEPP_STATE_PLACE", "1")
    monkeypatch.setenv("WEPP_DELTA_PLACE", delta)
    arena, reads, _ = synth.config_shape("C2", scale=0.02)
    _check_state_path(arena, reads, path=path)


def test_rescore_tile_and_generic_kernels_agree_at_size():
    """Beyond what the oracle finishes quickly: 200k nodes, 60k reads, 700 candidates — the tile kernel over the
    resident reads against the generic one-thread-per-read kernel of wepp_rescore_reads (itself oracle-checked)."""
    arena, reads, _ = synth.config_shape("C2", scale=0.2)
    reads = reads.slice(0, 60000)
    rng = np.random.default_rng(8)
    cand = rng.choice(arena.n_nodes, 700, replace=False).astype(np.int32)
    p = Placer(0)
    p.set_arena(arena)
    p.set_reads(reads)
    md, _, off, idx = p.rescore(cand)
    gmd, _, goff, gidx = p.rescore_reads(reads, cand)
    assert np.array_equal(md, gmd) and np.array_equal(off, goff) and np.array_equal(idx, gidx)
    md2, _, off2, _ = p.rescore(cand, want_argmin="count")
    assert np.array_equal(md2, gmd) and np.array_equal(off2, goff)
    p.close()


@pytest.mark.parametrize("seed,n_nodes,n_reads", [(41, 1500, 500), (42, 3000, 900)])
def test_peak_loop_matches_restated_filter(seed, n_nodes, n_reads):
    """wepp_filter_peaks (GPU-driven peak loop) vs oracle/peaks.py, itself pinned on the reference's object code."""
    from oracle import peaks
    arena, reads = cases.small_case(seed=seed, n_nodes=n_nodes, n_reads=n_reads)
    rng = np.random.default_rng(seed)
    leaf_count = rng.integers(1, 6, arena.n_nodes).astype(np.int32)
    ids = ["n%d" % v for v in range(arena.n_nodes)]
    order = sorted(range(arena.n_nodes), key=lambda i: ids[i])
    rank = np.zeros(arena.n_nodes, np.int32)
    rank[order] = np.arange(arena.n_nodes, dtype=np.int32)
    opk, onb = peaks.filter_peaks(arena, reads, leaf_count, ids)
    p = Placer(0)
    p.set_arena(arena)
    p.set_reads(reads)
    pk, nb = p.filter_peaks(leaf_count, rank)
    assert np.array_equal(pk, opk)
    assert np.array_equal(nb, onb)
    p.close()


def test_full_size_properties():
    """Size-independent properties at a size the oracle cannot finish: (1) degrees are
    conserved — sum over nodes of counts[v][b] == sum over reads in bin b of degree*multiplicity;
    (2) sum of scores == sum of degree/(1+p) over reads with a non-empty EPP set; (3) idempotence."""
    arena, reads, _ = synth.config_shape("C2", scale=0.2)   # 200k nodes x 200k reads
    p = Placer(0)
    p.set_arena(arena)
    p.set_reads(reads)
    p.place(0, 0)
    mp, mu = p.read_results()
    sc, ct = p.node_results()
    bins = np.minimum(reads.start // (arena.genome_size // 50), 49)
    expect = np.bincount(bins, weights=reads.degree.astype(np.float64) * mu, minlength=50)
    assert np.array_equal(ct.sum(axis=0, dtype=np.int64), expect.astype(np.int64))
    tot = (reads.degree / (1.0 + mp))[mu > 0].sum()
    assert abs(sc.sum() - tot) <= 1e-9 * tot
    assert mp.min() >= 0 and mu.min() >= 1 and mu.max() <= arena.n_nodes
    p.place(0, 0)
    mp2, mu2 = p.read_results()
    sc2, ct2 = p.node_results()
    assert np.array_equal(mp, mp2) and np.array_equal(mu, mu2) and np.array_equal(ct, ct2)
    np.testing.assert_allclose(sc, sc2, rtol=1e-12)
    p.close()


def test_allreduce_hook_single_rank_is_identity():
    """wepp_set_allreduce with a hook that sums over ONE rank (nothing to add): same results, and the library asks for
    the cell histogram + true counts in set_reads and for the accumulators in place (the multi-rank run is
    tests/peer_worker.py, which needs >= 2 GPUs)."""
    arena, reads = cases.small_case(seed=23, n_reads=4000)
    o = oracle.cartesian_map(arena, reads, None, n_threads=4)
    p = Placer(0)
    p.set_arena(arena)
    seen = []
    p.set_allreduce(lambda ptr, count, dtype, stream: seen.append((count, dtype)) or 0)
    p.set_reads(reads)
    assert [d for _, d in seen] == [0, 1]           # int32 histogram, int64 true counts
    p.place(0, 0)
    assert [d for _, d in seen[2:]] == [2, 0]       # float64 weights, int32 degrees
    assert p.stats()["place_path"] == 2
    mp, mu = p.read_results()
    sc, ct = p.node_results()
    assert np.array_equal(mp, o["max_parsimony"]) and np.array_equal(mu, o["multiplicity"]) and np.array_equal(ct, o["counts"])
    np.testing.assert_allclose(sc, o["score"], rtol=SCORE_RTOL, atol=1e-15)
    p.set_allreduce(None)
    with pytest.raises(Exception):
        p.place(0, 0)                                # the plan has to be derived again
    p.close()


@pytest.mark.parametrize("name", ["tiny1", "star", "small"])
def test_state_verify_mode(name, monkeypatch, capfd):
    """WEPP_STATE_VERIFY=1: the distinct states are told apart by two 64-bit hashes + size (probabilistically exact);
    the audit walk compares every evaluated list entry's actual state with the entries stored for its state."""
    monkeypatch.setenv("WEPP_STATE_VERIFY", "1")
    arena, reads = _rescore_case(name)
    _check_state_path(arena, reads)
    if name != "tiny1":   # (tiny trees may overflow the states' position cap and never build states)
        assert "every entry equals its state's representative" in capfd.readouterr().err


@pytest.mark.gpu
@pytest.mark.parametrize("ranks", [1, 2, 4])
@pytest.mark.parametrize("name,q", [("small", 16), ("star", 32), ("tiny1", 8)])
def test_group_place_matches_oracle(name, q, ranks):
    """wepp_group_*: reads dealt round-robin over the ranks of one process, one plan agreed through the group's own
    peer-memory all-reduce (cell histogram, true counts, per-(bucket, state) accumulators): per-read results gathered
    in the caller's order and the merged per-node results equal the oracle's over the whole read set — on every rank."""
    from wepp_b200.multigpu import Group
    arena, reads = _rescore_case(name)
    o = oracle.cartesian_map(arena, reads, None, n_threads=4)
    grp = Group([0] * ranks)
    grp.set_arena(arena)
    for _ in range(2):   # a second read set through the same group: events and barriers are reusable
        grp.set_reads(reads)
        grp.place()
        mp, mu = grp.read_results()
        assert np.array_equal(mp, o["max_parsimony"]) and np.array_equal(mu, o["multiplicity"])
        for r in range(ranks):
            sc, ct = grp.node_results(r)
            assert np.array_equal(ct, o["counts"]), r
            np.testing.assert_allclose(sc, o["score"], rtol=1e-9, atol=1e-15)
    grp.close()


@pytest.mark.gpu
def test_group_reports_a_failing_rank():
    """a rank that fails outside an exchange must not leave the others waiting at the barrier"""
    from wepp_b200._lib import WeppError
    from wepp_b200.multigpu import Group
    arena, reads = _rescore_case("small")
    grp = Group([0, 0])
    with pytest.raises(WeppError):
        grp.set_reads(reads)          # no arena yet: every rank fails
    grp.set_arena(arena)
    grp.set_reads(reads)
    grp.place()
    grp.close()
