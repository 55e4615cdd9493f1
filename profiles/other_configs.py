#!/usr/bin/env python
"""The other BASELINE.json configurations at their named shapes, on one GPU: they are parity-test cases, not bench
lines (bench.py measures C3), so this only records that they run at full size, how long a placement takes, and the
size-independent checks (degrees conserved, score mass, idempotence).
usage: python profiles/other_configs.py C2|C4 [read_scale]   -> one JSON line"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wepp_b200 import synth                      # noqa: E402
from wepp_b200.placement import Placer           # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "C4"
    rscale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    g = 29903
    n = {"C2": 1_000_000, "C3": 8_000_000, "C4": 8_000_000, "MID": 8_000_000, "FRAG": 8_000_000}[name]
    r = int({"C2": 1_000_000, "C3": 10_000_000, "C4": 1_000_000, "MID": 1_000_000, "FRAG": 1_000_000}[name] * rscale)
    arena = synth.make_arena(n, g, synth.SEED)
    if name == "C4":
        reads = synth.make_reads(arena, r, synth.SEED, amplicons=synth.amplicon_scheme(g, 29, 1058, 1201, synth.SEED),
                                 full_amplicon=True, err=0.03, n_rate=0.05)
    elif name == "FRAG":  # trimmed short reads: starts anywhere, lengths 40-150 (not a BASELINE config: bucket fragmentation)
        reads = synth.make_reads(arena, r, synth.SEED, amplicons=synth.amplicon_scheme(g, 5000, 60, 150, synth.SEED),
                                 full_amplicon=True)
    elif name == "MID":   # 450-550-base windows (not a BASELINE config: sizes the reads-per-lane heuristic)
        reads = synth.make_reads(arena, r, synth.SEED, amplicons=synth.amplicon_scheme(g, 60, 450, 550, synth.SEED),
                                 full_amplicon=True, err=0.01, n_rate=0.02)
    else:
        reads = synth.make_reads(arena, r, synth.SEED)
    p = Placer(0, stripe_width=int(os.environ.get("WEPP_STRIPE_WIDTH", "16")), reads_per_lane=int(os.environ.get("WEPP_READS_PER_LANE", "0")))
    t0 = time.perf_counter(); p.set_arena(arena); t_arena = time.perf_counter() - t0
    t0 = time.perf_counter(); p.set_reads(reads); p.lib.wepp_sync(p.h); t_reads = time.perf_counter() - t0
    t0 = time.perf_counter(); p.set_reads(reads); p.lib.wepp_sync(p.h); t_reads2 = time.perf_counter() - t0
    for _ in range(2):
        p.place(0, 0)
    st = p.stats()
    mp, mu = p.read_results()
    sc, ct = p.node_results()
    bins = np.minimum(reads.start // (g // 50), 49)
    expect = np.bincount(bins, weights=reads.degree.astype(np.float64) * mu, minlength=50).astype(np.int64)
    ok_counts = bool(np.array_equal(ct.sum(axis=0, dtype=np.int64), expect))
    tot = float((reads.degree / (1.0 + mp))[mu > 0].sum())
    ok_score = bool(abs(float(sc.sum()) - tot) <= 1e-9 * tot)
    p.place(0, 0)
    mp2, mu2 = p.read_results()
    out = {"config": name, "nodes": arena.n_nodes, "reads": reads.n_reads, "mean_window": float((reads.end - reads.start + 1).mean()),
           "set_arena_s": t_arena, "set_reads_first_s": t_reads, "set_reads_again_s": t_reads2,
           "place_kernel_ms": st["ms_scan_kernel"], "node_kernels_ms": st["ms_node_kernels"],
           "reads_per_s_kernel": reads.n_reads / ((st["ms_scan_kernel"] + st["ms_node_kernels"]) / 1e3),
           "reads_per_tile": st["reads_per_tile"], "n_tiles": st["n_tiles"], "n_lists": st["n_lists"],
           "scanned_entries": st["scanned_entries"], "degrees_conserved": ok_counts, "score_mass_ok": ok_score,
           "idempotent": bool(np.array_equal(mp, mp2) and np.array_equal(mu, mu2)),
           "max_parsimony_mean": float(mp.mean()), "multiplicity_median": float(np.median(mu))}
    print(json.dumps(out), flush=True)
    p.close()


if __name__ == "__main__":
    main()
