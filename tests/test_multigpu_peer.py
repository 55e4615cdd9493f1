"""The multi-GPU exchange step over peer memory (wepp_peer_merge) on real GPUs: needs >= 2 devices, so it is
skipped on single-GPU boxes (the host-side sharding logic is covered by tests/test_dist_gloo.py on CPU)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_peer_merge_matches_oracle_and_allreduce():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "peer_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "PEER_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_group_on_distinct_devices_matches_oracle():
    """wepp_group over real peer memory: one rank per GPU (2, 4 or 8 of them), the in-process exchange kernel reading
    and writing the other devices' buffers over NVLink."""
    import numpy as np
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    import oracle
    from wepp_b200 import synth
    from wepp_b200.multigpu import Group
    world = 8 if n >= 8 else (4 if n >= 4 else 2)
    arena = synth.make_arena(20000, 29903, 3)
    reads = synth.make_reads(arena, 20000, 3)
    o = oracle.cartesian_map(arena, reads, None, n_threads=8)
    grp = Group(list(range(world)))
    grp.set_arena(arena)
    grp.set_reads(reads)
    grp.place()
    mp, mu = grp.read_results()
    assert np.array_equal(mp, o["max_parsimony"]) and np.array_equal(mu, o["multiplicity"])
    sc0, ct0 = grp.node_results(0)
    assert np.array_equal(ct0, o["counts"])
    np.testing.assert_allclose(sc0, o["score"], rtol=1e-9, atol=1e-15)
    for r in range(1, world):
        sc, ct = grp.node_results(r)
        assert np.array_equal(ct, ct0) and np.array_equal(sc, sc0)
    grp.close()
