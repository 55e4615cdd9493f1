#!/usr/bin/env python
"""wepp_group at the bench shape: N ranks of one process (one per GPU), 1.25 M reads per rank against the 8 M-node tree,
exchanges by the library's peer-memory kernel.  Reports per-step device times (max over ranks) and checks the merged
results against a single-GPU placement of the same reads.  usage: python profiles/group_scale.py N [scale]"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
from wepp_b200 import synth, _lib
from wepp_b200.multigpu import Group
from wepp_b200.placement import Placer

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
n_nodes, per = max(int(8_000_000 * scale), 1000), max(int(1_250_000 * scale), 256)
arena = synth.make_arena(n_nodes, 29903, synth.SEED)
reads = synth.make_reads(arena, per * N, synth.SEED)
g = Group(list(range(N)))
t0 = time.time(); g.set_arena(arena); t_arena = time.time() - t0
t0 = time.time(); g.set_reads(reads); t_reads = time.time() - t0
g.place()
lib = _lib.load()
steps = []
for _ in range(6):
    t0 = time.perf_counter()
    g.place()
    wall = (time.perf_counter() - t0) * 1e3
    st = []
    for r in range(N):
        s = _lib.WeppStats()
        _lib.check(lib.wepp_get_stats(lib.wepp_group_handle(g.g, r), C.byref(s)))
        st.append(s.as_dict())
    steps.append({"wall_ms": round(wall, 3), "place_ms_max": round(max(s["ms_place_total"] for s in st), 3),
                  "scan_ms_max": round(max(s["ms_scan_kernel"] for s in st), 3), "exchange_ms_max": round(max(s["ms_exchange"] for s in st), 3),
                  "node_ms_max": round(max(s["ms_node_kernels"] for s in st), 3), "path": st[0]["place_path"]})
mp, mu = g.read_results()
sc, ct = g.node_results(0)
same = all(np.array_equal(g.node_results(r)[0], sc) for r in range(1, N))
out = {"n_gpus": N, "nodes": n_nodes, "reads_total": per * N, "set_arena_s": round(t_arena, 3), "set_reads_s": round(t_reads, 3),
       "steps": steps[1:], "reads_per_s": per * N / (min(s["wall_ms"] for s in steps[1:]) * 1e-3), "ranks_bit_identical": bool(same)}
g.close()
if "--check" in sys.argv:
    p = Placer(0); p.set_arena(arena); p.set_reads(reads); p.place(0, 0)
    mp1, mu1 = p.read_results(); sc1, ct1 = p.node_results(); p.close()
    out["vs_one_gpu"] = {"max_parsimony": bool(np.array_equal(mp, mp1)), "multiplicity": bool(np.array_equal(mu, mu1)),
                         "counts": bool(np.array_equal(ct, ct1)), "score_max_rel": float(np.max(np.abs(sc - sc1) / np.maximum(np.abs(sc1), 1e-300) * (sc1 != 0)))}
print(json.dumps(out))
