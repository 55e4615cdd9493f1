"""Read sharding and the per-node exchange step of the multi-GPU placement driver.

Reads are independent units (the reference's own decomposition, src/WEPP/initial_filter.cpp:152):
rank r places the contiguous slice shard_bounds(R, r, world); per-read outputs stay local.  The
per-node arrays are combined with one sum all-reduce — the GPU form of the reference's
mutex-serialised chunk merge (initial_filter.cpp:199-211).  The functions take torch tensors on
any device, so the same code runs over NCCL (GPU buffers of the library, see bench.py) and over
gloo (CPU tests)."""
from __future__ import annotations


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
