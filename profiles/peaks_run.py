#!/usr/bin/env python
"""wepp_filter_peaks (the whole initial filter: cartesian_map + greedy peak loop + neighbour expansion,
initial_filter.cpp:455-506) at the C3 shard shape on one GPU: wall time through the C ABI.
usage: python profiles/peaks_run.py [scale] [--gpus N]   (N > 1: wepp_group, the reads dealt over N GPUs of one process)"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                      # noqa: E402
from wepp_b200.placement import Placer            # noqa: E402


def main():
    scale = float(sys.argv[1]) if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else 1.0
    gpus = int(sys.argv[sys.argv.index("--gpus") + 1]) if "--gpus" in sys.argv else 1
    arena, reads = bench.workload("C3", scale, 0)
    n = arena.n_nodes
    rng = np.random.default_rng(3)
    is_leaf = np.ones(n, bool)
    is_leaf[arena.parent[1:]] = False
    leaf_count = np.where(is_leaf, rng.integers(1, 4, n), 0).astype(np.int32)
    id_rank = rng.permutation(n).astype(np.int32)
    if gpus > 1:
        from wepp_b200.multigpu import Group
        p = Group(list(range(gpus)))
    else:
        p = Placer(0)
    p.set_arena(arena)
    p.set_reads(reads)
    p.place() if gpus > 1 else p.place(0, 0)       # warm-up
    t0 = time.perf_counter()
    peaks, nbrs = p.filter_peaks(leaf_count, id_rank)
    dt = time.perf_counter() - t0
    print(json.dumps({"nodes": n, "reads": reads.n_reads, "gpus": gpus, "filter_peaks_s": dt, "n_peaks": int(len(peaks)),
                      "n_with_neighbours": int(len(nbrs)), "peaks_head": [int(x) for x in peaks[:8]]}), flush=True)
    p.close()


if __name__ == "__main__":
    main()
