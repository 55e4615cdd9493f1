"""World-size-2 gloo test of the multi-GPU host logic (read sharding + per-node all-reduce +
per-read gather), with the CPU oracle standing in for the per-rank placement."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from tests import cases
from wepp_b200 import multigpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    arena, reads = cases.small_case(seed=23, n_nodes=800, n_reads=301)
    lo, hi = multigpu.shard_bounds(reads.n_reads, rank, world)
    o = oracle.cartesian_map(arena, reads.slice(lo, hi), None)
    score = torch.from_numpy(o["score"].copy())
    counts = torch.from_numpy(o["counts"].reshape(-1).copy())
    multigpu.allreduce_node_arrays(score, counts)
    mp_all = multigpu.gather_read_results(torch.from_numpy(o["max_parsimony"].copy()), reads.n_reads, rank, world)
    mu_all = multigpu.gather_read_results(torch.from_numpy(o["multiplicity"].copy()), reads.n_reads, rank, world)
    if rank == 0:
        np.savez(os.path.join(out_dir, "r0.npz"), score=score.numpy(), counts=counts.numpy(), mp=mp_all.numpy(),
                 mu=mu_all.numpy())
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "r0.npz")
    arena, reads = cases.small_case(seed=23, n_nodes=800, n_reads=301)
    o = oracle.cartesian_map(arena, reads, None)
    assert np.array_equal(got["mp"], o["max_parsimony"]) and np.array_equal(got["mu"], o["multiplicity"])
    assert np.array_equal(got["counts"].reshape(-1, 50), o["counts"])
    np.testing.assert_allclose(got["score"], o["score"], rtol=1e-12)


def test_shard_bounds_partition():
    for n in (0, 1, 7, 1000, 1_250_001):
        for w in (1, 2, 4, 8):
            b = [multigpu.shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
