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
