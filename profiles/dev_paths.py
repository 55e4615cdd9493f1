#!/usr/bin/env python
"""Development aid: the three placement paths of wepp_place(0, 0) on one shard — sparse corrections over the states
(delta_place.cuh), the dense state kernel (state_place.cuh), the Euler-list scan (place_kernel) — cross-checked
against each other at the given scale, with per-path timings.
usage: python profiles/dev_paths.py [scale] [--euler]"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wepp_b200 import synth
from wepp_b200.placement import Placer

scale = float(sys.argv[1]) if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else 1.0
n_nodes, n_reads = max(int(8_000_000 * scale), 1000), max(int(1_250_000 * scale), 256)
arena = synth.make_arena(n_nodes, 29903, synth.SEED)
reads = synth.make_reads(arena, n_reads, synth.SEED)
out = {}
res = {}
paths = [("delta", {"WEPP_DELTA_PLACE": "1"}), ("states", {"WEPP_DELTA_PLACE": "0"})]
if "--euler" in sys.argv:
    paths.append(("euler", {"WEPP_STATE_PLACE": "0"}))
p = Placer(0)
p.set_arena(arena)
for name, env in paths:
    for k in ("WEPP_DELTA_PLACE", "WEPP_STATE_PLACE"):
        os.environ.pop(k, None)
    os.environ.update(env)
    t0 = time.time()
    p.set_reads(reads)
    p.place(0, 0)
    p.sync()
    first = time.time() - t0
    ms = []
    for _ in range(5):
        p.place(0, 0)
        st = p.stats()
        ms.append((st["ms_scan_kernel"], st["ms_node_kernels"]))
    mp, mu = p.read_results()
    sc, ct = p.node_results()
    res[name] = (mp, mu, sc, ct.sum(axis=0, dtype=np.int64), ct[:: max(1, n_nodes // 100000)].copy())
    out[name] = {"first_call_s": round(first, 3), "ms_scan": [round(a, 3) for a, _ in ms], "ms_node": [round(b, 3) for _, b in ms],
                 "path": st["place_path"], "n_states": st["n_states"], "n_window_groups": st["n_window_groups"],
                 "n_lists": st["n_lists"], "n_tiles": st["n_tiles"]}
    print(name, json.dumps(out[name]), flush=True)
ref = res[paths[0][0]]
for name, _ in paths[1:]:
    r = res[name]
    ok = {"max_parsimony": bool(np.array_equal(ref[0], r[0])), "multiplicity": bool(np.array_equal(ref[1], r[1])),
          "counts_colsum": bool(np.array_equal(ref[3], r[3])), "counts_sample": bool(np.array_equal(ref[4], r[4])),
          "score_max_rel": float(np.max(np.abs(ref[2] - r[2]) / np.maximum(np.abs(r[2]), 1e-300) * (r[2] != 0))),
          "score_zero_pattern": bool(np.array_equal(ref[2] == 0, r[2] == 0))}
    print("delta vs", name, json.dumps(ok), flush=True)
p.close()
