import os, sys, time, numpy as np
sys.path.insert(0, "/root/repo")
os.environ["WEPP_RESCORE_TIMING"] = "1"
import bench
from wepp_b200.placement import Placer
arena, reads = bench.workload(1.0, 0)
p = Placer(0, stripe_width=int(os.environ.get("Q", "32")))
p.set_arena(arena); p.set_reads(reads)
rng = np.random.default_rng(1)
pool = rng.choice(arena.n_nodes, 5000, replace=False).astype(np.int32)
for i in range(3):
    t0 = time.perf_counter(); p.rescore(pool, want_argmin=False); print("total ms", (time.perf_counter() - t0) * 1e3, file=sys.stderr)
