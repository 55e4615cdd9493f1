#!/usr/bin/env python
"""Development aid: wall time of every C-ABI call of an end-to-end step (set_reads / place / results out)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wepp_b200 import synth
from wepp_b200.placement import Placer
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
arena, reads, _ = synth.config_shape("C3", scale=scale, n_reads=max(int(1_250_000 * scale), 256))
p = Placer(0)
p.set_arena(arena)
for it in range(4):
    t = [time.perf_counter()]
    p.set_reads(reads); t.append(time.perf_counter())
    p.set_mapped(None); t.append(time.perf_counter())
    p.place(0, 0); t.append(time.perf_counter())
    p.read_results(); t.append(time.perf_counter())
    p.node_summary(); t.append(time.perf_counter())
    print("iter", it, " ".join(f"{n}={1e3 * (b - a):.2f}ms" for n, a, b in zip(("set_reads", "set_mapped", "place", "read_results", "node_summary"), t, t[1:])), flush=True)
p.close()
