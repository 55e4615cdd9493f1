"""torchrun worker of tests/test_multigpu_peer.py: every rank places its read shard on its own GPU, the per-node
arrays are merged (a) by the peer-memory kernel and (b) by the NCCL all-reduce, and rank 0 checks both against
the oracle run over ALL reads."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import oracle                                    # noqa: E402
from tests import cases                          # noqa: E402
from wepp_b200 import multigpu                   # noqa: E402
from wepp_b200.placement import Placer           # noqa: E402


def cuda_view(ptr, n, typestr, device):
    class _V:
        pass
    v = _V()
    v.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}
    return torch.as_tensor(v, device=f"cuda:{device}")


def main():
    rank, world, dev = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl")
    for seed, n_nodes, n_reads in [(23, 800, 301), (31, 5000, 1500), (37, 64, 40)]:
        arena, reads = cases.small_case(seed=seed, n_nodes=n_nodes, n_reads=n_reads)
        lo, hi = multigpu.shard_bounds(reads.n_reads, rank, world)
        p = Placer(dev)
        p.set_stream(torch.cuda.current_stream().cuda_stream)
        p.set_arena(arena)
        p.set_reads(reads.slice(lo, hi))
        peer = multigpu.PeerMerge(p, rank, world, dev)
        for _ in range(2):                       # twice: the barriers must make the second place safe
            p.place(0, 0, sync=False)
            peer.merge()
        sc, dv = p.node_summary()
        # (b) the all-reduce path on a fresh placement
        p.place(0, 0, sync=False)
        sp, sb = p.device_buffer(1)
        cp, cb = p.device_buffer(2)
        multigpu.allreduce_node_arrays(cuda_view(sp, sb // 8, "<f8", dev), cuda_view(cp, cb // 4, "<i4", dev))
        torch.cuda.synchronize()
        sc2, ct2 = p.node_results()
        o = oracle.cartesian_map(arena, reads, None, n_threads=2)
        bins = np.minimum(reads.start // (arena.genome_size // 50), 49)
        true_counts = np.bincount(bins, weights=reads.degree.astype(np.float64), minlength=50)
        with np.errstate(divide="ignore", invalid="ignore"):
            prop = o["counts"] / true_counts[None, :]
        odv = (prop > 0.005).sum(axis=1) / float((true_counts != 0).sum())
        np.testing.assert_allclose(sc, o["score"], rtol=1e-9, atol=1e-15)
        assert np.array_equal(dv, odv), f"rank {rank}: dist_divergence differs"
        assert np.array_equal(ct2, o["counts"])
        np.testing.assert_allclose(sc2, o["score"], rtol=1e-9, atol=1e-15)
        peer.close()
        p.close()
        # (c) one plan for all ranks + exchange of the per-(bucket, state) accumulators (wepp_set_allreduce): every
        #     rank ends up with the merged per-node arrays; per-read results are the rank's own slice
        for env in ({}, {"WEPP_DELTA_PLACE": "2"}, {"WEPP_DELTA_PLACE": "0"}, {"WEPP_STATE_PLACE": "0"}):
            for k in ("WEPP_DELTA_PLACE", "WEPP_STATE_PLACE"):
                os.environ.pop(k, None)
            os.environ.update(env)
            q = Placer(dev)
            q.set_stream(torch.cuda.current_stream().cuda_stream)
            q.set_arena(arena)
            shared = multigpu.SharedPlan(q, dev)
            for _ in range(2):
                q.set_reads(reads.slice(lo, hi))
                q.place(0, 0, sync=False)
            torch.cuda.synchronize()
            assert shared.calls >= 8, shared.calls
            sc3, ct3 = q.node_results()
            sc4, dv4 = q.node_summary()
            mp3, mu3 = q.read_results()
            assert np.array_equal(ct3, o["counts"]), f"rank {rank} {env}: counts differ"
            np.testing.assert_allclose(sc3, o["score"], rtol=1e-9, atol=1e-15)
            np.testing.assert_allclose(sc4, o["score"], rtol=1e-9, atol=1e-15)
            assert np.array_equal(dv4, odv), f"rank {rank} {env}: dist_divergence differs"
            assert np.array_equal(mp3, o["max_parsimony"][lo:hi]) and np.array_equal(mu3, o["multiplicity"][lo:hi])
            shared.close()
            q.close()
        for k in ("WEPP_DELTA_PLACE", "WEPP_STATE_PLACE"):
            os.environ.pop(k, None)
    dist.barrier()
    if rank == 0:
        print("PEER_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
