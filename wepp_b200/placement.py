"""Host-side mirror of the reference's placement interface over the C ABI.

`WeppFilter` keeps the reference's names (src/WEPP/initial_filter.hpp:14-65): after
`cartesian_map()` it holds `max_parismony` [sic], `parsimony_multiplicity`, the EPP cache and
the per-node `score` / `mapped_read_counts`, exactly the state `wepp_filter` and `haplotype`
carry after src/WEPP/initial_filter.cpp:139-239.  All compute happens in libwepp_b200.so on
the GPU; this module only marshals numpy arrays.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, ptr

NUM_RANGE_BINS = 50          # src/WEPP/config.hpp:13
MAX_CACHED_EPP_SIZE = 2048   # src/WEPP/config.hpp:9
READ_DIST_FACTOR_THRESHOLD = 0.5 / 100  # src/WEPP/config.hpp:18


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


class Placer:
    """Thin object wrapper of one wepp_handle (one GPU)."""

    def __init__(self, device: int = 0, stripe_width: int | None = None, reads_per_lane: int | None = None):
        self.lib = _lib.load()
        h = C.c_void_p()
        check(self.lib.wepp_create(int(device), C.byref(h)))
        self.h = h
        self.n_nodes = 0
        self.n_reads = 0
        if stripe_width is not None or reads_per_lane is not None:
            check(self.lib.wepp_set_options(self.h, int(stripe_width or 16), int(reads_per_lane or 0)))

    def close(self):
        if getattr(self, "h", None):
            self.lib.wepp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- inputs ------------------------------------------------------------------------------
    def set_arena(self, arena) -> None:
        self._arena = (_c(arena.parent, np.int32), _c(arena.mut_off, np.int64), _c(arena.mut_pos, np.int32),
                       _c(arena.mut_ref, np.uint8), _c(arena.mut_nuc, np.uint8))
        p, o, mp, mr, mn = self._arena
        check(self.lib.wepp_set_arena(self.h, p.shape[0], ptr(p), ptr(o), ptr(mp), ptr(mr), ptr(mn),
                                      int(arena.genome_size)))
        self.n_nodes = int(p.shape[0])
        self.genome_size = int(arena.genome_size)

    def set_reads(self, reads) -> None:
        r = (_c(reads.start, np.int32), _c(reads.end, np.int32), _c(reads.degree, np.int32),
             _c(reads.rm_off, np.int64), _c(reads.rm_pos, np.int32), _c(reads.rm_nuc, np.uint8))
        check(self.lib.wepp_set_reads(self.h, r[0].shape[0], *[ptr(x) for x in r]))
        self.n_reads = int(r[0].shape[0])
        self.reads_generation = getattr(self, "reads_generation", 0) + 1

    def set_mapped(self, mapped) -> None:
        m = None if mapped is None else _c(mapped, np.uint8)
        check(self.lib.wepp_set_mapped(self.h, ptr(m)))

    def set_allreduce(self, fn) -> None:
        """Read-sharded ranks (wepp_set_allreduce, include/wepp_b200.h): fn(dev_ptr: int, count: int, dtype: int,
        cuda_stream: int) -> 0 sums the device array in place over the ranks (dtype: 0 int32, 1 int64, 2 float64).
        None = single-rank behaviour.  Call before set_reads."""
        if fn is None:
            self._allreduce_cb = None
            check(self.lib.wepp_set_allreduce(self.h, None, None))
            return
        proto = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p)

        def tramp(_user, dev_ptr, count, dtype, stream):
            try:
                return int(fn(int(dev_ptr or 0), int(count), int(dtype), int(stream or 0)) or 0)
            except Exception:   # never let an exception cross the C frame
                import traceback
                traceback.print_exc()
                return 1
        self._allreduce_cb = proto(tramp)   # keep the thunk alive as long as the handle uses it
        check(self.lib.wepp_set_allreduce(self.h, C.cast(self._allreduce_cb, C.c_void_p), None))

    # -- compute -----------------------------------------------------------------------------
    def set_stream(self, cuda_stream: int) -> None:
        """Run all later work on the caller's CUDA stream (e.g. torch.cuda.current_stream().cuda_stream)."""
        check(self.lib.wepp_set_stream(self.h, C.c_void_p(int(cuda_stream))))

    def sync(self) -> None:
        check(self.lib.wepp_sync(self.h))

    def place(self, epp_cap: int = 0, epp_capacity: int = 0, sync: bool = True) -> None:
        check(self.lib.wepp_place(self.h, int(epp_cap), int(epp_capacity)))
        if sync:
            self.sync()

    def place_subset(self, read_idx, epp_cap: int = 0, epp_capacity: int = 0) -> None:
        idx = _c(read_idx, np.int64)
        check(self.lib.wepp_place_subset(self.h, idx.shape[0], ptr(idx), int(epp_cap), int(epp_capacity)))
        self.sync()

    # -- outputs -----------------------------------------------------------------------------
    def read_results(self):
        mp = np.empty(self.n_reads, np.int32)
        mu = np.empty(self.n_reads, np.int32)
        check(self.lib.wepp_get_read_results(self.h, ptr(mp), ptr(mu)))
        return mp, mu

    def node_results(self, want_counts: bool = True):
        sc = np.empty(self.n_nodes, np.float64)
        ct = np.empty((self.n_nodes, NUM_RANGE_BINS), np.int32) if want_counts else None
        check(self.lib.wepp_get_node_results(self.h, ptr(sc), ptr(ct)))
        return sc, ct

    def node_summary(self):
        """score[N] and dist_divergence[N] (initial_filter.cpp:214-231), the divergence computed on the device."""
        sc = np.empty(self.n_nodes, np.float64)
        dv = np.empty(self.n_nodes, np.float64)
        check(self.lib.wepp_get_node_summary(self.h, ptr(sc), ptr(dv)))
        return sc, dv

    def filter_peaks(self, leaf_count, id_rank):
        """wepp_filter::filter (initial_filter.cpp:455-506): returns (peaks, neighbours) as arena indices."""
        lc, ir = _c(leaf_count, np.int32), _c(id_rank, np.int32)
        out = np.empty(self.n_nodes, np.int32)
        npk, nout = C.c_int32(0), C.c_int32(0)
        check(self.lib.wepp_filter_peaks(self.h, ptr(lc), ptr(ir), ptr(out), out.shape[0], C.byref(npk), C.byref(nout)))
        return out[: npk.value].copy(), out[npk.value: nout.value].copy()

    def epp(self):
        off = np.empty(self.n_reads + 1, np.int64)
        n = C.c_int64(0)
        check(self.lib.wepp_get_epp(self.h, ptr(off), None, 0, C.byref(n)))
        nodes = np.empty(max(int(n.value), 1), np.int32)
        check(self.lib.wepp_get_epp(self.h, ptr(off), ptr(nodes), nodes.shape[0], C.byref(n)))
        return off, nodes[: int(n.value)]

    def cartesian_map_host(self, reads, mapped=None):
        """wepp_cartesian_map: host buffers in, host buffers out, everything inside one call."""
        r = (_c(reads.start, np.int32), _c(reads.end, np.int32), _c(reads.degree, np.int32),
             _c(reads.rm_off, np.int64), _c(reads.rm_pos, np.int32), _c(reads.rm_nuc, np.uint8))
        m = None if mapped is None else _c(mapped, np.uint8)
        n = r[0].shape[0]
        mp = np.empty(n, np.int32)
        mu = np.empty(n, np.int32)
        sc = np.empty(self.n_nodes, np.float64)
        ct = np.empty((self.n_nodes, NUM_RANGE_BINS), np.int32)
        check(self.lib.wepp_cartesian_map(self.h, n, *[ptr(x) for x in r], ptr(m), ptr(mp), ptr(mu), ptr(sc), ptr(ct)))
        self.n_reads = int(n)
        return mp, mu, sc, ct

    def rescore(self, cand_nodes, want_dist: bool = False, want_argmin=True):
        """wepp_rescore over the resident reads.  want_argmin: True = CSR argmin lists, "count" = only their
        offsets (am_off; the number of argmins per read), False = min distance only."""
        cand = _c(cand_nodes, np.int32)
        md = np.empty(self.n_reads, np.int32)
        dist = np.empty((self.n_reads, cand.shape[0]), np.int32) if want_dist else None
        off = np.empty(self.n_reads + 1, np.int64) if want_argmin else None
        lists = want_argmin is True
        cap = self.n_reads * cand.shape[0] if lists else 0
        idx = np.empty(max(cap, 1), np.int32) if lists else None
        check(self.lib.wepp_rescore(self.h, cand.shape[0], ptr(cand), ptr(md), ptr(dist), ptr(off), ptr(idx), cap))
        if lists:
            idx = idx[: int(off[-1])]
        return md, dist, off, idx

    def rescore_reads(self, reads, cand_nodes, want_dist: bool = False, want_argmin: bool = True):
        """wepp_rescore_reads: the same over reads handed in directly (generic one-thread-per-read kernel)."""
        cand = _c(cand_nodes, np.int32)
        r = (_c(reads.start, np.int32), _c(reads.end, np.int32), _c(reads.rm_off, np.int64), _c(reads.rm_pos, np.int32),
             _c(reads.rm_nuc, np.uint8))
        n = r[0].shape[0]
        md = np.empty(n, np.int32)
        dist = np.empty((n, cand.shape[0]), np.int32) if want_dist else None
        off = np.empty(n + 1, np.int64) if want_argmin else None
        cap = n * cand.shape[0] if want_argmin else 0
        idx = np.empty(max(cap, 1), np.int32) if want_argmin else None
        check(self.lib.wepp_rescore_reads(self.h, n, *[ptr(x) for x in r], cand.shape[0], ptr(cand), ptr(md), ptr(dist),
                                          ptr(off), ptr(idx), cap))
        if want_argmin:
            idx = idx[: int(off[-1])]
        return md, dist, off, idx

    PEER_BLOB_BYTES = 664

    def peer_export(self) -> bytes:
        """wepp_peer_export: this rank's CUDA IPC handles + per-bin read counts, to be exchanged between ranks."""
        buf = C.create_string_buffer(self.PEER_BLOB_BYTES)
        check(self.lib.wepp_peer_export(self.h, buf))
        return buf.raw

    def peer_open(self, rank: int, world: int, blobs: bytes) -> None:
        assert len(blobs) == world * self.PEER_BLOB_BYTES
        check(self.lib.wepp_peer_open(self.h, rank, world, C.c_char_p(blobs)))

    def peer_merge(self) -> None:
        check(self.lib.wepp_peer_merge(self.h))

    def peer_close(self) -> None:
        check(self.lib.wepp_peer_close(self.h))

    def device_buffer(self, which: int):
        p = C.c_void_p()
        n = C.c_int64()
        check(self.lib.wepp_device_buffer(self.h, which, C.byref(p), C.byref(n)))
        return int(p.value), int(n.value)

    def stats(self) -> dict:
        st = _lib.WeppStats()
        check(self.lib.wepp_get_stats(self.h, C.byref(st)))
        return st.as_dict()


class WeppFilter:
    """Mirror of `wepp_filter` (src/WEPP/initial_filter.hpp:14-65) for the cartesian_map stage."""

    def __init__(self, placer: Placer):
        self.placer = placer
        self.max_parismony = None            # spelling as in the reference
        self.parsimony_multiplicity = None
        self.epp_positions_cache = None      # (offsets, nodes) CSR, empty range = not cached
        self.score = None
        self.mapped_read_counts = None
        self.dist_divergence = None

    def cartesian_map(self, reads, mapped=None, epp_capacity: int | None = None):
        p = self.placer
        p.set_reads(reads)
        p.set_mapped(mapped)
        cap = int(epp_capacity if epp_capacity is not None else min(reads.n_reads * MAX_CACHED_EPP_SIZE + 1024, 1 << 28))
        p.place(MAX_CACHED_EPP_SIZE, cap)
        self.max_parismony, self.parsimony_multiplicity = p.read_results()
        self.epp_positions_cache = p.epp()
        self.score, self.mapped_read_counts = p.node_results()
        # dist_divergence, src/WEPP/initial_filter.cpp:214-231
        g = p.genome_size
        bin_size = g // NUM_RANGE_BINS
        bins = np.minimum(np.asarray(reads.start) // bin_size, NUM_RANGE_BINS - 1)
        true_counts = np.bincount(bins, weights=np.asarray(reads.degree, dtype=np.float64),
                                  minlength=NUM_RANGE_BINS).astype(np.int64)
        active = int((true_counts != 0).sum())
        with np.errstate(divide="ignore", invalid="ignore"):
            prop = self.mapped_read_counts / true_counts[None, :].astype(np.float64)
        self.dist_divergence = (prop > READ_DIST_FACTOR_THRESHOLD).sum(axis=1) / max(active, 1)
        return self


def build_arena(tree, reads, masked=None):
    """Host-side arena construction (wepp_arena_build): returns (Arena, Reads, info) where Arena is
    the flattened condensed tree, Reads the masked reads and info carries `source`, `leaf_count`
    and the folded-node CSR (`map_off`, `map_nodes`)."""
    from .synth import Arena, Reads
    lib = _lib.load()
    a = (_c(tree.parent, np.int32), _c(tree.mut_off, np.int64), _c(tree.mut_pos, np.int32), _c(tree.mut_ref, np.uint8),
         _c(tree.mut_nuc, np.uint8))
    m = _c(masked if masked is not None else [], np.int32)
    r = (_c(reads.start, np.int32), _c(reads.end, np.int32), _c(reads.rm_off, np.int64), _c(reads.rm_pos, np.int32),
         _c(reads.rm_nuc, np.uint8))
    h = C.c_void_p()
    check(lib.wepp_arena_build(a[0].shape[0], *[ptr(x) for x in a], int(tree.genome_size), m.shape[0], ptr(m),
                               r[0].shape[0], *[ptr(x) for x in r], C.byref(h)))
    try:
        n, nm, nrm, nmap = C.c_int32(), C.c_int64(), C.c_int64(), C.c_int64()
        check(lib.wepp_arena_dims(h, C.byref(n), C.byref(nm), C.byref(nrm), C.byref(nmap)))
        n, nm, nrm, nmap = n.value, nm.value, nrm.value, nmap.value
        out = {"parent": np.zeros(n, np.int32), "source": np.zeros(n, np.int32), "leaf_count": np.zeros(n, np.int32),
               "mut_off": np.zeros(n + 1, np.int64), "mut_pos": np.zeros(nm, np.int32), "mut_ref": np.zeros(nm, np.uint8),
               "mut_nuc": np.zeros(nm, np.uint8), "map_off": np.zeros(n + 1, np.int64), "map_nodes": np.zeros(nmap, np.int32)}
        check(lib.wepp_arena_get(h, *[ptr(out[k]) for k in ("parent", "source", "leaf_count", "mut_off", "mut_pos",
                                                           "mut_ref", "mut_nuc", "map_off", "map_nodes")]))
        ro, rp, rn = np.zeros(reads.n_reads + 1, np.int64), np.zeros(nrm, np.int32), np.zeros(nrm, np.uint8)
        check(lib.wepp_arena_get_reads(h, ptr(ro), ptr(rp), ptr(rn)))
    finally:
        lib.wepp_arena_free(h)
    arena = Arena(int(tree.genome_size), tree.ref_codes, out["parent"], out["mut_off"], out["mut_pos"], out["mut_ref"],
                  out["mut_nuc"])
    mreads = Reads(np.asarray(reads.start, np.int32), np.asarray(reads.end, np.int32), np.asarray(reads.degree, np.int32),
                   ro, rp, rn, dict(getattr(reads, "meta", {})))
    return arena, mreads, out
