"""Read sharding and the per-node exchange step of the multi-GPU placement driver.

Reads are independent units (the reference's own decomposition, src/WEPP/initial_filter.cpp:152):
rank r places the contiguous slice shard_bounds(R, r, world); per-read outputs stay local.  The
per-node arrays are combined with one sum all-reduce — the GPU form of the reference's
mutex-serialised chunk merge (initial_filter.cpp:199-211).  The functions take torch tensors on
any device, so the same code runs over NCCL (GPU buffers of the library, see bench.py) and over
gloo (CPU tests)."""
from __future__ import annotations

import numpy as np


def shard_bounds(n_reads: int, rank: int, world: int) -> tuple[int, int]:
    return n_reads * rank // world, n_reads * (rank + 1) // world


def allreduce_node_arrays(score, counts, group=None) -> None:
    """In-place sum over ranks of score[N] (float64) and counts[N*50] (int32)."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(score, op=dist.ReduceOp.SUM, group=group)


_TYPESTR = {0: ("<i4", 4), 1: ("<i8", 8), 2: ("<f8", 8)}


def cuda_view(ptr: int, n: int, typestr: str, device: int):
    """torch tensor over a library-owned device buffer."""
    import torch

    class _V:
        pass
    v = _V()
    v.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}
    return torch.as_tensor(v, device=f"cuda:{device}")


class SharedPlan:
    """Read-sharded ranks that agree on ONE plan and exchange the per-(bucket, state) accumulators of a placement
    instead of the per-node arrays (wepp_set_allreduce): the library asks for an in-place sum over the ranks of (1) the
    (window, bin) cell histogram + true read counts in set_reads and (2) the accumulators in place; this class lends it
    torch.distributed.all_reduce (NCCL) on the placer's stream.  ~100 MB per step at 8 M nodes against 1.66 GB of
    per-node arrays; every rank then holds the merged score / read counts / dist_divergence.
    bucket_bytes: large arrays go out in slices of this size so that NCCL pipelines them (0 = one call)."""

    def __init__(self, placer, device: int, group=None, stream=None):
        import torch
        self.placer, self.device, self.group = placer, device, group
        self.calls = 0
        self.bytes = 0
        self._stream = stream
        placer.set_allreduce(self._allreduce)

    def _allreduce(self, dev_ptr: int, count: int, dtype: int, cuda_stream: int) -> int:
        import torch
        import torch.distributed as dist
        if count <= 0:
            return 0
        typestr, size = _TYPESTR[dtype]
        t = cuda_view(dev_ptr, count, typestr, self.device)
        ext = torch.cuda.ExternalStream(cuda_stream, device=self.device) if cuda_stream else torch.cuda.current_stream(self.device)
        with torch.cuda.stream(ext):
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        self.calls += 1
        self.bytes += count * size
        return 0

    def close(self) -> None:
        self.placer.set_allreduce(None)


class PeerMerge:
    """The exchange step over NVLink peer memory (wepp_peer_*, include/wepp_b200.h): every rank sums its slice of
    the nodes straight out of the other ranks' HBM, evaluates dist_divergence on the merged rows and stores the
    merged score / dist_divergence into every rank.  torch.distributed only moves the 664-byte IPC blobs once
    and provides the two stream-ordered barriers around the kernel (a 4-byte all-reduce on the placement stream)."""

    def __init__(self, placer, rank: int, world: int, device: int, group=None):
        import torch
        import torch.distributed as dist
        self.placer, self.group, self.rank, self.world = placer, group, rank, world
        self._open()
        self.token = torch.zeros(1, dtype=torch.int32, device=f"cuda:{device}")

    def _open(self) -> None:
        """Publish this rank's buffers and true read counts, map the peers'.  The summed true read counts belong to
        the read sets resident at this moment: after a set_reads the library refuses to merge until this ran again."""
        import torch.distributed as dist
        blobs = [None] * self.world
        dist.all_gather_object(blobs, self.placer.peer_export(), group=self.group)
        self.placer.peer_open(self.rank, self.world, b"".join(blobs))
        self._generation = getattr(self.placer, "reads_generation", 0)

    def merge(self) -> None:
        """Enqueue: barrier, merge kernel, barrier — all on the current (= the placer's) stream."""
        import torch.distributed as dist
        if getattr(self.placer, "reads_generation", 0) != self._generation:   # every rank changed its reads: exchange again
            self._open()
        dist.all_reduce(self.token, group=self.group)   # every rank's placement + scans are done
        self.placer.peer_merge()
        dist.all_reduce(self.token, group=self.group)   # every rank is done reading: the next place may overwrite

    def close(self) -> None:
        self.placer.peer_close()


def gather_read_results(local, n_reads: int, rank: int, world: int, group=None):
    """Concatenate per-read int32 results of all ranks in read order (rank slices are contiguous)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or world == 1:
        return local
    sizes = [shard_bounds(n_reads, r, world)[1] - shard_bounds(n_reads, r, world)[0] for r in range(world)]
    m = max(sizes)                                    # all_gather wants equal shapes: pad, then trim
    padded = torch.zeros(m, dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    bufs = [torch.empty(m, dtype=local.dtype, device=local.device) for _ in sizes]
    dist.all_gather(bufs, padded, group=group)
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)])


class Group:
    """wepp_group (include/wepp_b200.h): several GPUs of one box driven from THIS process — a handle and a host
    thread per rank inside the library, reads dealt round-robin, exchanges by the library's own peer-memory kernel
    (no torch.distributed, no NCCL).  `devices` may repeat a device (ranks sharing one GPU)."""

    def __init__(self, devices):
        import ctypes as C
        from . import _lib
        self.lib = _lib.load()
        self._check = _lib.check
        dev = np.ascontiguousarray(devices, np.int32)
        g = C.c_void_p()
        self._check(self.lib.wepp_group_create(int(dev.shape[0]), _lib.ptr(dev), C.byref(g)))
        self.g = g
        self.n_ranks = int(dev.shape[0])
        self.n_nodes = 0
        self.n_reads = 0

    def close(self):
        if getattr(self, "g", None):
            self.lib.wepp_group_destroy(self.g)
            self.g = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_arena(self, arena):
        from ._lib import ptr
        a = (np.ascontiguousarray(arena.parent, np.int32), np.ascontiguousarray(arena.mut_off, np.int64),
             np.ascontiguousarray(arena.mut_pos, np.int32), np.ascontiguousarray(arena.mut_ref, np.uint8),
             np.ascontiguousarray(arena.mut_nuc, np.uint8))
        self._check(self.lib.wepp_group_set_arena(self.g, a[0].shape[0], *[ptr(x) for x in a], int(arena.genome_size)))
        self.n_nodes = int(a[0].shape[0])

    def set_reads(self, reads):
        from ._lib import ptr
        r = (np.ascontiguousarray(reads.start, np.int32), np.ascontiguousarray(reads.end, np.int32),
             np.ascontiguousarray(reads.degree, np.int32), np.ascontiguousarray(reads.rm_off, np.int64),
             np.ascontiguousarray(reads.rm_pos, np.int32), np.ascontiguousarray(reads.rm_nuc, np.uint8))
        self._check(self.lib.wepp_group_set_reads(self.g, r[0].shape[0], *[ptr(x) for x in r]))
        self.n_reads = int(r[0].shape[0])

    def place(self):
        self._check(self.lib.wepp_group_place(self.g))

    def read_results(self):
        from ._lib import ptr
        mp, mu = np.empty(self.n_reads, np.int32), np.empty(self.n_reads, np.int32)
        self._check(self.lib.wepp_group_get_read_results(self.g, ptr(mp), ptr(mu)))
        return mp, mu

    def node_results(self, rank: int = 0):
        """the merged per-node score and mapped_read_counts as rank `rank` holds them (identical on every rank)"""
        from ._lib import ptr
        from .placement import NUM_RANGE_BINS
        sc = np.empty(self.n_nodes, np.float64)
        ct = np.empty((self.n_nodes, NUM_RANGE_BINS), np.int32)
        self._check(self.lib.wepp_get_node_results(self.lib.wepp_group_handle(self.g, int(rank)), ptr(sc), ptr(ct)))
        return sc, ct

    def node_summary(self, rank: int = 0):
        from ._lib import ptr
        sc, dv = np.empty(self.n_nodes, np.float64), np.empty(self.n_nodes, np.float64)
        self._check(self.lib.wepp_get_node_summary(self.lib.wepp_group_handle(self.g, int(rank)), ptr(sc), ptr(dv)))
        return sc, dv

    def filter_peaks(self, leaf_count, id_rank):
        import ctypes as C
        from ._lib import ptr
        lc, ir = np.ascontiguousarray(leaf_count, np.int32), np.ascontiguousarray(id_rank, np.int32)
        out = np.empty(self.n_nodes, np.int32)
        npk, nout = C.c_int32(0), C.c_int32(0)
        self._check(self.lib.wepp_group_filter_peaks(self.g, ptr(lc), ptr(ir), ptr(out), out.shape[0], C.byref(npk), C.byref(nout)))
        return out[: npk.value].copy(), out[npk.value: nout.value].copy()
