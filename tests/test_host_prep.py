"""CPU suite: the product's HOST logic (Euler stripes with signed-delta tables, read plan)
checked against the oracle through a numpy emulation of the device algorithm."""
import ctypes as C
import re
import os

import numpy as np
import pytest

import oracle
from tests import cases, emulate
from wepp_b200 import synth
from wepp_b200 import _lib
from wepp_b200._lib import ptr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("seed,q", [(0, 4), (1, 8), (2, 16), (3, 32), (4, 5), (5, 64), (6, 1), (7, 8)])
def test_emulated_scan_matches_oracle_tiny(seed, q):
    arena, reads, mapped = cases.tiny_case(seed)
    for m in (None, mapped):
        o = oracle.cartesian_map(arena, reads, m, epp_cap=arena.n_nodes)
        best, mult, epps = emulate.emulate_place(arena, reads, m, q=q)
        assert np.array_equal(best, o["max_parsimony"])
        assert np.array_equal(mult, o["multiplicity"])
        for r in range(reads.n_reads):
            assert np.array_equal(epps[r], o["epp_nodes"][o["epp_off"][r]:o["epp_off"][r + 1]])


def test_emulated_scan_matches_oracle_small():
    arena, reads = cases.small_case()
    o = oracle.cartesian_map(arena, reads, None, n_threads=4, epp_cap=arena.n_nodes)
    best, mult, epps = emulate.emulate_place(arena, reads, None, q=32)
    assert np.array_equal(best, o["max_parsimony"])
    assert np.array_equal(mult, o["multiplicity"])
    for r in range(0, reads.n_reads, 13):
        assert np.array_equal(epps[r], o["epp_nodes"][o["epp_off"][r]:o["epp_off"][r + 1]])


def test_emulated_scan_matches_oracle_star_tree():
    arena, reads, mapped = cases.star_case(1)
    for m in (None, mapped):
        o = oracle.cartesian_map(arena, reads, m, n_threads=4, epp_cap=arena.n_nodes)
        best, mult, epps = emulate.emulate_place(arena, reads, m, q=16)
        assert np.array_equal(best, o["max_parsimony"])
        assert np.array_equal(mult, o["multiplicity"])
        for r in range(0, reads.n_reads, 7):
            assert np.array_equal(epps[r], o["epp_nodes"][o["epp_off"][r]:o["epp_off"][r + 1]])


def test_leaf_events_are_single_point_entries():
    """A leaf's event is one POINT entry (key bit 0 set) at the leaf's index; an internal node's event is
    an ENTER/EXIT boundary pair."""
    arena, _, _ = cases.star_case(2)
    ent, off = emulate.host_stripes(arena, 16)
    n = arena.n_nodes
    children = np.bincount(arena.parent[1:], minlength=n)
    point = (ent[:, 0] & 1).astype(bool)
    idx = (ent[:, 0] >> 1).astype(np.int64)
    assert np.all(children[idx[point]] == 0)
    n_leaf_events = int(np.diff(arena.mut_off)[children == 0].sum())
    assert int(point.sum()) <= n_leaf_events and int(point.sum()) > 0.9 * n_leaf_events
    assert ent.shape[0] < 2 * int(arena.mut_off[-1])


def test_stripes_are_grouped_and_sorted():
    arena, _ = cases.small_case()
    ent, off = emulate.host_stripes(arena, 32)
    assert off[0] == 0 and off[-1] == ent.shape[0]
    for s in range(off.shape[0] - 1):
        e = ent[off[s]:off[s + 1]]
        assert np.all(e[:, 1] // 32 == s)
        assert np.all(np.diff(e[:, 0].astype(np.int64)) >= 0)


def test_read_plan_is_a_permutation_and_buckets_cover_windows():
    arena, reads = cases.small_case()
    lib = _lib.load()
    r = reads.n_reads
    perm = np.full(r, -1, np.int64)
    qs = np.zeros(r, np.int32); qe = np.zeros(r, np.int32); bn = np.zeros(r, np.int32)
    rpt = C.c_int32(0)
    nt = _lib.check(lib.wepp_host_read_plan(arena.genome_size, 32, 0, r, ptr(reads.start), ptr(reads.end),
                                            ptr(reads.degree), ptr(reads.rm_off), ptr(reads.rm_pos), ptr(reads.rm_nuc),
                                            ptr(perm), ptr(qs), ptr(qe), ptr(bn), C.byref(rpt)))
    assert nt > 0 and rpt.value in (64, 128, 256)
    assert np.array_equal(np.sort(perm), np.arange(r))
    s, e = reads.start[perm], reads.end[perm]
    assert np.all(qs * 32 <= s) and np.all(e <= qe * 32 + 31)
    assert np.array_equal(bn, np.minimum(s // (arena.genome_size // 50), 49))


def _host_plan(arena, reads, q):
    lib = _lib.load()
    r = reads.n_reads
    perm = np.full(r, -1, np.int64)
    qs = np.zeros(r, np.int32); qe = np.zeros(r, np.int32); bn = np.zeros(r, np.int32)
    rpt = C.c_int32(0)
    nt = _lib.check(lib.wepp_host_read_plan(arena.genome_size, q, 0, r, ptr(reads.start), ptr(reads.end),
                                            ptr(reads.degree), ptr(reads.rm_off), ptr(reads.rm_pos), ptr(reads.rm_nuc),
                                            ptr(perm), ptr(qs), ptr(qe), ptr(bn), C.byref(rpt)))
    return nt, rpt.value, perm, qs, qe, bn


def test_bucket_coarsening_keeps_windows_covered_and_saves_tile_entries(monkeypatch):
    """Reads with scattered starts and lengths (trimmed short reads): sparse (stripe range, bin) buckets are merged
    into wider ones of the same first stripe.  Every window must stay inside its list's range, the count bin must be
    the read's own, and the tile-entries (tiles x list width) must not grow."""
    arena, _ = cases.small_case(seed=13, n_reads=10)
    rng = np.random.default_rng(4)
    r = 4000
    start = rng.integers(1, arena.genome_size - 160, r).astype(np.int32)
    end = (start + rng.integers(30, 150, r)).astype(np.int32)
    reads = synth.Reads(start, end, np.ones(r, np.int32), np.zeros(r + 1, np.int64), np.zeros(0, np.int32), np.zeros(0, np.uint8))
    q = 8

    def cost(plan):
        nt, rpt, perm, qs, qe, bn = plan
        s, e = reads.start[perm], reads.end[perm]
        assert np.array_equal(np.sort(perm), np.arange(r))
        assert np.all(qs * q <= s) and np.all(e <= qe * q + q - 1)
        assert np.array_equal(bn, np.minimum(s // (arena.genome_size // 50), 49))
        key = np.stack([qs, qe, bn], axis=1)
        uniq, counts = np.unique(key, axis=0, return_counts=True)
        tiles = -(-counts // rpt)
        return int((tiles * (uniq[:, 1] - uniq[:, 0] + 1)).sum()), len(uniq)

    merged = cost(_host_plan(arena, reads, q))
    monkeypatch.setenv("WEPP_NO_BUCKET_MERGE", "1")
    plain = cost(_host_plan(arena, reads, q))
    assert merged[1] < plain[1] and merged[0] <= plain[0]


@pytest.mark.parametrize("bad", ["parent", "dup_pos", "pos_range", "read_nuc", "read_order", "read_window"])
def test_malformed_inputs_are_rejected(bad):
    arena, reads, _ = cases.tiny_case(1)
    lib = _lib.load()
    a = [arena.parent.copy(), arena.mut_off.copy(), arena.mut_pos.copy(), arena.mut_ref.copy(), arena.mut_nuc.copy()]
    rd = [reads.start.copy(), reads.end.copy(), reads.degree.copy(), reads.rm_off.copy(), reads.rm_pos.copy(), reads.rm_nuc.copy()]
    if bad == "parent":
        a[0][5] = 7
    elif bad == "dup_pos":
        v = int(np.flatnonzero(np.diff(a[1]) >= 2)[0]); a[2][a[1][v] + 1] = a[2][a[1][v]]
    elif bad == "pos_range":
        a[2][0] = arena.genome_size + 1
    elif bad == "read_nuc":
        rd[5][0] = 5
    elif bad == "read_order":
        k = int(np.flatnonzero(np.diff(rd[3]) >= 2)[0]); o = rd[3][k]; rd[4][o], rd[4][o + 1] = rd[4][o + 1], rd[4][o]
    elif bad == "read_window":
        rd[1][0] = arena.genome_size + 5
    if bad in ("parent", "dup_pos", "pos_range"):
        rc = lib.wepp_host_euler_stripes(arena.n_nodes, *[ptr(x) for x in a], arena.genome_size, 8, None, 0, None, 0)
    else:
        rc = lib.wepp_host_read_plan(arena.genome_size, 8, 0, reads.n_reads, *[ptr(x) for x in rd], None, None, None, None, None)
    assert rc == -1
    assert len(lib.wepp_last_error()) > 0


def test_library_exports_every_declared_symbol():
    """The C-ABI library loads (no GPU needed) and exports exactly what include/wepp_b200.h declares."""
    hdr = open(os.path.join(ROOT, "include", "wepp_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(wepp_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = C.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/wepp_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), "python binding table out of sync with the header"
    assert lib.wepp_abi_version() == 1


def test_product_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from wepp_b200.placement import Placer
    with pytest.raises(_lib.WeppError) as ei:
        Placer(0)
    assert "no CPU fallback" in str(ei.value)


def test_group_fails_loudly_without_gpu_and_checks_its_arguments():
    """wepp_group_*: no device, no group (and no crash); argument checks need no device at all."""
    import torch
    lib = _lib.load()
    g = C.c_void_p()
    assert lib.wepp_group_create(0, None, C.byref(g)) == -1 and b"ranks" in lib.wepp_last_error()
    assert lib.wepp_group_create(17, None, C.byref(g)) == -1
    assert lib.wepp_group_size(None) == 0 and lib.wepp_group_handle(None, 0) is None and lib.wepp_group_take(None, 0) is None
    assert lib.wepp_group_place(None) == -1 and lib.wepp_group_run(None, None, None) == -1
    lib.wepp_group_destroy(None)
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from wepp_b200.multigpu import Group
    with pytest.raises(_lib.WeppError) as ei:
        Group([0, 0])
    assert "no CPU fallback" in str(ei.value)
