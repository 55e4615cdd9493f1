#!/usr/bin/env python
"""bench.py — reads placed/s of WEPP's parsimonious read placement on B200(s).

    python bench.py --gpus N --steps K --warmup W              # our CUDA path
    python bench.py --impl reference --gpus N --steps K ...    # the reference's CPU placement
    python bench.py --config C2|C4|C1 ...                      # the other named shapes (parity-test cases, same JSON)

Workload (BASELINE.json configs[2], the one the metric is quoted on): a synthetic public-scale SARS-CoV-2 MAT
(8M arena nodes, genome 29,903) x 150-bp reads at the ends of the 99 ARTIC v4.1 amplicons, read-sharded: every GPU
places 1.25M collapsed reads against the replicated tree (weak scaling; 8 GPUs = the full 10M reads).  One step =
one cartesian_map over the rank's shard: placement kernel (delta_place_kernel: sparse corrections per read over the
distinct window-restricted haplotypes of its window) + node_tile_kernel (per-node score / read counts / divergence
bin count in one pass), and for N>1 the exchange step (one kernel over NVLink peer memory, or --exchange nccl: the
all-reduce of the per-node score / read-count arrays).

One JSON line is printed by rank 0 (see README / DESIGN.md for the keys).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GENOME = 29903
# per-GPU shard of each named shape; C3 is the bench line, the others are parity-test cases
CONFIGS = {
    "C3": {"reads_per_gpu": 1_250_000, "what": "C3 synthetic SARS-CoV-2-scale MAT x ARTIC v4.1 150-bp collapsed reads (read-sharded; 8 GPUs = 10M reads)"},
    "C2": {"reads_per_gpu": 1_000_000, "what": "C2 SARS-CoV-2 quick-start shape: 1M-node MAT x 1M ARTIC v4.1 150-bp collapsed reads"},
    "C4": {"reads_per_gpu": 1_000_000, "what": "C4 synthetic ONT sample: 8M-node MAT x 1M full-amplicon reads on the midnight scheme (1.1-1.2 kb windows)"},
    "C1": {"reads_per_gpu": 200_000, "what": "C1 RSV-A quick-start shape: 50k-node MAT (genome 15,222) x 200k 150-bp collapsed reads on the RSV-A scheme"},
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C3", choices=sorted(CONFIGS))
    ap.add_argument("--scale", type=float, default=1.0, help="shrink nodes and reads (development only)")
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="CPU-baseline budget per measurement")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the untimed cross-checks of the step's results")
    ap.add_argument("--exchange", default="states", choices=["states", "peer", "nccl"],
                    help="N>1 exchange step: states = one plan on all ranks, all-reduce of the per-(bucket, state) accumulators "
                         "(default); peer = peer-memory merge kernel over the per-node arrays; nccl = all-reduce of the per-node arrays")
    ap.add_argument("--no-c5", action="store_true", help="skip the candidate re-scoring (C5) measurement")
    ap.add_argument("--c5-generic", action="store_true", help="also time the generic K4 kernel once (slow)")
    return ap.parse_args()


def workload(config: str, scale: float, rank: int):
    from wepp_b200 import synth
    n_reads = max(int(CONFIGS[config]["reads_per_gpu"] * scale), 256)
    arena, reads, _ = synth.config_shape(config, scale=max(scale, 1000 / 8_000_000), n_reads=n_reads,
                                         read_seed=synth.SEED + 1000 * rank)
    return arena, reads


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.device), "-lms", "100"], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                c = [x.strip() for x in line.split(",")]
                if len(c) < 9:
                    continue
                try:
                    sm.append(float(c[1])); mx.append(float(c[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


def cuda_view(ptr: int, n: int, typestr: str, device: int):
    """torch tensor over a library-owned device buffer (for NCCL)."""
    import torch

    class _V:
        pass
    v = _V()
    v.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}
    return torch.as_tensor(v, device=f"cuda:{device}")


# ---------------------------------------------------------------------------------------------
def _marginal(run, n_max: int, seconds: float, cores: int, floor: int):
    """reads/s of `run(n) -> seconds` as a MARGINAL rate between two sample sizes (n, 2n), sized so that the pair
    takes about `seconds`: fixed costs (thread start-up, page faults of the per-thread arrays) cancel."""
    n = max(cores, min(n_max // 2, floor))
    dt = run(n)
    while dt < seconds / 6.0 and 2 * n <= n_max // 2:   # grow until a pair (n, 2n) fills the budget
        n = int(min(n_max // 2, max(2 * n, n * (seconds / 3.0) / max(dt, 1e-6))))
        n = max(cores, n // cores * cores)
        dt = run(n)
    n2 = min(n_max, 2 * n)
    dt2 = run(n2)
    rate = (n2 - n) / max(dt2 - dt, 1e-9) if n2 > n and dt2 > dt else n2 / dt2
    return rate, n, dt, n2, dt2


def cpu_reference(arena, reads, seconds: float):
    """The reference's CPU placement of the step's workload on the box's host cores, all threads.

    value: the CPU restatement of the reference's algorithm WITH the reference's range trees (oracle/wepp_oracle.cpp:
    single_read_tree over the per-range compressed trees of arena.cpp:68-169, built from the whole shard's read windows
    as the reference builds them), timed on its read loop (the reference's own "cartesian mapping took" boundary,
    initial_filter.cpp:144,238) on a bounded sample, as a marginal rate between two sample sizes.  The reference's own
    object code (oracle/_ref) cannot hold the 8M-node bench tree in a bounded sample (its arena keeps O(N x depth)
    stack_muts, ~5 KB per node, and cartesian_map allocates N x 208 B per chunk, initial_filter.cpp:155-158), so
    `reference_own_code` times it next to the port on a smaller tree of the same generator: the port is the FASTER of
    the two (no std::string per mutation, no per-chunk dense arrays), i.e. the baseline reported here flatters the CPU."""
    import oracle
    from wepp_b200 import synth
    cores = os.cpu_count() or 1

    def run_port(n):
        o = oracle.cartesian_map(arena, reads.slice(0, n), None, n_threads=cores, want_node=False, range_trees=True,
                                 range_reads=reads)
        return max(o["seconds_map"], 1e-6)

    rate, n1, t1, n2, t2 = _marginal(run_port, reads.n_reads, seconds, cores, floor=cores * 64)
    out = {"value": rate, "unit": "reads/s", "cores": cores, "kind": "port",
           "sample": f"marginal rate between {n1} and {n2} of the step's {reads.n_reads} reads vs all {arena.n_nodes} nodes "
                     f"({t1:.1f} s and {t2:.1f} s of read loop; range trees from the whole shard's windows, built untimed)",
           "algorithm": "restated single_read_tree over the reference's range trees (arena.cpp:68-169)"}
    # the reference's own object code next to the port, on a tree both can hold
    try:
        from oracle import ref as oref
        if oref.available() and seconds >= 5.0:
            n_small = int(min(arena.n_nodes, 250_000))
            sm_arena = synth.make_arena(n_small, arena.genome_size, synth.SEED)
            sm_reads = synth.make_reads(sm_arena, 65_536, synth.SEED, amplicons=synth.primer_scheme("ARTICv4_1"))
            g = sm_arena.genome_size
            cs = np.arange(1, g + 1, 150, dtype=np.int32)
            ce = np.minimum(cs + 149, g).astype(np.int32)
            cover = synth.Reads(np.concatenate([sm_reads.start, cs]), np.concatenate([sm_reads.end, ce]),
                                np.concatenate([sm_reads.degree, np.ones(cs.size, np.int32)]),
                                np.concatenate([sm_reads.rm_off, np.full(cs.size, sm_reads.rm_off[-1], np.int64)]),
                                sm_reads.rm_pos, sm_reads.rm_nuc)
            t0 = time.perf_counter()
            sess = oref.Session(sm_arena, cover, threads=cores)
            t_build = time.perf_counter() - t0
            if sess.n_nodes == sm_arena.n_nodes:
                def run_ref(n):
                    return max(sess.cartesian_map(n_sel=n, want_node=False, want_epp=False)["ms"] / 1e3, 1e-6)

                def run_port_small(n):
                    o = oracle.cartesian_map(sm_arena, sm_reads.slice(0, n), None, n_threads=cores, want_node=False,
                                             range_trees=True, range_reads=sm_reads)
                    return max(o["seconds_map"], 1e-6)

                r_ref, a1, _, a2, _ = _marginal(run_ref, sm_reads.n_reads, seconds / 2.0, cores, floor=cores * 64)
                r_port, _, _, _, _ = _marginal(run_port_small, sm_reads.n_reads, seconds / 4.0, cores, floor=cores * 64)
                out["reference_own_code"] = {
                    "nodes": n_small, "reference_reads_per_s": r_ref, "port_reads_per_s": r_port, "port_over_reference": r_port / r_ref,
                    "sample": f"oracle/_ref (the reference's translation units, shim-compiled) vs the port on a {n_small}-node "
                              f"tree of the same generator, marginal rates ({a1} -> {a2} reads for the reference; its arena "
                              f"build, {t_build:.1f} s, untimed)"}
            sess.close()
    except Exception as e:   # the reference library is optional test infrastructure
        out["reference_own_code"] = {"unavailable": repr(e)[:200]}
    return out


def run_reference(args, rank: int):
    if rank != 0:
        return
    arena, reads = workload(args.config, args.scale, 0)
    base = cpu_reference(arena, reads, args.cpu_seconds)
    v = float(base["value"])
    out = {"impl": "reference", "metric": "reads placed/s", "value": v, "unit": "reads/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": reads.n_reads / v * 1e3, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
           "config": config_dict(arena, reads, args), "cpu_baseline": base,
           "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def config_dict(arena, reads, args):
    scheme = {"C1": "RSVA_all_primers_best_hits", "C4": "midnight"}.get(args.config, "ARTICv4_1")
    return {"workload": f"{CONFIGS[args.config]['what']}: {arena.n_nodes} arena nodes, {arena.n_events} events, genome "
                        f"{arena.genome_size}, {reads.n_reads} reads per GPU",
            "name": args.config, "nodes": arena.n_nodes, "reads_per_gpu": reads.n_reads, "seed": 20260101,
            "amplicons": f"wepp_b200/data/{scheme}.amplicons.tsv (from the reference's primers/{scheme}.bed)",
            "l2": "inputs exceed L2 (posting lists 170 MB + per-window base scores + 1.6 GB per-node arrays are re-streamed every step)"}


# ---------------------------------------------------------------------------------------------
def run_ours(args, rank: int, world: int, local_rank: int):
    import torch
    import torch.distributed as dist
    from wepp_b200 import multigpu, synth
    from wepp_b200.placement import Placer

    torch.cuda.set_device(local_rank)
    dev = local_rank
    arena, reads = workload(args.config, args.scale, rank)
    k_env = int(os.environ.get("WEPP_READS_PER_LANE", "0"))   # development knob
    q_env = int(os.environ.get("WEPP_STRIPE_WIDTH", "16"))
    p = Placer(dev, stripe_width=q_env, reads_per_lane=k_env)
    stream = torch.cuda.current_stream()
    p.set_stream(stream.cuda_stream)
    t0 = time.perf_counter()
    p.set_arena(arena)
    t_arena = time.perf_counter() - t0
    shared = None
    if world > 1 and args.exchange == "states":
        shared = multigpu.SharedPlan(p, dev)   # the library all-reduces the cell histogram (set_reads) and the accumulators (place)
    t0 = time.perf_counter()
    p.set_reads(reads)
    t_reads = time.perf_counter() - t0

    def allreduce_nodes():
        if world == 1:
            return
        sp, sb = p.device_buffer(1)
        cp, cb = p.device_buffer(2)
        multigpu.allreduce_node_arrays(cuda_view(sp, sb // 8, "<f8", dev), cuda_view(cp, cb // 4, "<i4", dev))

    # N>1 exchange step: the peer-memory merge kernel (sums the ranks' per-node arrays over NVLink and evaluates
    # dist_divergence in the same pass) or, with --exchange nccl, the all-reduce of score[N] and counts[N][50]
    peer = None
    if world > 1 and args.exchange == "peer":
        peer = multigpu.PeerMerge(p, rank, world, dev)

    def exchange(full_counts: bool = False):
        if world == 1 or shared is not None:   # shared plan: wepp_place exchanges the accumulators itself
            return
        if peer is not None and not full_counts:
            peer.merge()
        else:
            allreduce_nodes()

    def step():
        p.place(0, 0, sync=False)
        exchange()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(dev)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    scan_ms = []
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    # the dominant kernel's own launch time (CUDA events on the launching stream), separate short loop
    for _ in range(3):
        p.place(0, 0, sync=True)
        scan_ms.append(p.stats()["ms_scan_kernel"])
    st = p.stats()
    # the exchange step alone: ranks aligned by a barrier first, CUDA events around it on the launching stream
    exch_ms = None
    if world > 1 and shared is not None:
        t = torch.tensor([float(st["ms_exchange"])], device=f"cuda:{dev}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        exch_ms = float(t.item())
    elif world > 1:
        xs = []
        for _ in range(3):
            p.place(0, 0, sync=False)
            barrier()
            x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            x0.record()
            exchange()
            x1.record()
            torch.cuda.synchronize()
            xs.append(x0.elapsed_time(x1))
        t = torch.tensor([float(np.median(xs))], device=f"cuda:{dev}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        exch_ms = float(t.item())
    if world > 1:
        t = torch.tensor([ms], device=f"cuda:{dev}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    reads_total = reads.n_reads * world
    value = reads_total / (ms_per_step / 1e3)

    # ---- end to end through the C ABI with host buffers ----------------------------------------
    # One reference-facing cartesian_map per step: host reads in (packing, H2D, per-window Euler
    # lists), placement, NCCL merge, results out to pinned host memory.  "e2e" returns what the
    # reference's later stages read (per-read min parsimony / multiplicity, per-node score and
    # dist_divergence); "e2e_full_counts" also brings back the 200-byte-per-node
    # mapped_read_counts matrix that the reference keeps in haplotype but never reads again.
    e2e = e2e_full = None
    if not args.no_e2e:
        n, r = arena.n_nodes, reads.n_reads
        mp = torch.empty(r, dtype=torch.int32, pin_memory=True).numpy()
        mu = torch.empty(r, dtype=torch.int32, pin_memory=True).numpy()
        sc = torch.empty(n, dtype=torch.float64, pin_memory=True).numpy()
        dv = torch.empty(n, dtype=torch.float64, pin_memory=True).numpy()
        ct = torch.empty((n, 50), dtype=torch.int32, pin_memory=True).numpy()
        from wepp_b200._lib import check, ptr

        def pinned(a):
            return torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()

        # the step's inputs live in pinned host memory, in the caller's (unsorted) order
        reads_h = synth.Reads(pinned(reads.start), pinned(reads.end), pinned(reads.degree), pinned(reads.rm_off),
                              pinned(reads.rm_pos), pinned(reads.rm_nuc))

        def e2e_step(full: bool):
            p.set_reads(reads_h)        # H2D of the raw reads, device keying/bucketing, per-window Euler list build
            p.set_mapped(None)
            p.place(0, 0, sync=False)
            exchange(full_counts=full)
            check(p.lib.wepp_get_read_results(p.h, ptr(mp), ptr(mu)))
            if rank == 0 or world == 1:
                if full:
                    check(p.lib.wepp_get_node_results(p.h, ptr(sc), ptr(ct)))
                else:
                    check(p.lib.wepp_get_node_summary(p.h, ptr(sc), ptr(dv)))

        def time_e2e(full: bool):
            e2e_step(full)
            barrier()
            t0 = time.perf_counter()
            k_e2e = max(1, min(args.steps, 3))
            for _ in range(k_e2e):
                e2e_step(full)
            barrier()
            dt = (time.perf_counter() - t0) / k_e2e
            if world > 1:
                t = torch.tensor([dt], device=f"cuda:{dev}")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            return dt

        def e2e_phases():
            """One more e2e step with a synchronisation after each phase (not part of the e2e figure)."""
            ph = {}
            barrier()
            t0 = time.perf_counter()
            p.set_reads(reads_h)
            p.set_mapped(None)
            torch.cuda.synchronize()
            ph["reads_in_keying_lists_ms"] = (time.perf_counter() - t0) * 1e3
            t0 = time.perf_counter()
            p.place(0, 0, sync=False)
            exchange()
            torch.cuda.synchronize()
            ph["place_and_exchange_ms"] = (time.perf_counter() - t0) * 1e3
            t0 = time.perf_counter()
            check(p.lib.wepp_get_read_results(p.h, ptr(mp), ptr(mu)))
            ph["read_results_out_ms"] = (time.perf_counter() - t0) * 1e3
            t0 = time.perf_counter()
            check(p.lib.wepp_get_node_summary(p.h, ptr(sc), ptr(dv)))
            ph["node_summary_out_ms"] = (time.perf_counter() - t0) * 1e3
            barrier()
            return ph

        q = st["stripe_width"]
        n_stripes = GENOME // q + 1
        n_cells = n_stripes * min(n_stripes, 4000 // q + 2) * min(50, q // (GENOME // 50) + 2)   # keying histogram
        h2d = int(r * 12 + (r + 1) * 8 + reads.rm_pos.shape[0] * 5                     # the reads
                  + st["n_tiles"] * 16 + st["n_lists"] * 32 + st["n_buckets"] * 24 + n_cells * 4)   # descriptors
        d2h_keys = n_cells * 4 + 416
        dt = time_e2e(False)
        e2e = {"value": reads_total / dt, "unit": "reads/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": int(r * 8 + n * 16 + d2h_keys), "ms_per_step": dt * 1e3,
               "outputs": "max_parsimony, multiplicity per read; score, dist_divergence per node",
               "note": "every step re-uploads the reads, keys and buckets them, rebuilds the per-window Euler lists and the "
                       "read set's window groups (sort by window, base scores and histogram per window); the per-tree indices "
                       "(stripe rank table, distinct window-restricted haplotypes + posting lists of the window lists, tile "
                       "tables) are built on the first call and kept while read sets map to the same window lists and bins; "
                       "e2e_cold is one sample on a fresh handle with everything built",
               "phases": e2e_phases()}
        dt = time_e2e(True)
        e2e_full = {"value": reads_total / dt, "unit": "reads/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": int(r * 8 + n * (8 + 200) + d2h_keys), "ms_per_step": dt * 1e3,
                    "outputs": "as e2e, with mapped_read_counts[N][50] instead of dist_divergence"}

    parity = None
    x_bytes = 0
    if shared is not None:   # bytes one placement hands to the collective
        b0 = shared.bytes
        p.place(0, 0, sync=True)
        x_bytes = shared.bytes - b0
    if not args.no_parity and world == 1:
        parity = parity_block(p, arena, reads, dev)
    elif not args.no_parity:
        parity = multi_rank_checks(p, arena, reads, dev, world)
    e2e_cold = None
    if not args.no_e2e and rank == 0 and world == 1:
        e2e_cold = cold_sample(arena, reads, dev, q_env, k_env)

    if peer is not None:
        p.set_reads(reads)
        p.place(0, 0, sync=False)
        peer.merge()
        peer.close()
    c5 = None
    if not args.no_c5:
        c5 = run_c5(args, p, arena, reads, world, dev, barrier)

    if rank != 0:
        return
    peak, peak_src = peaks()
    k_ms = float(np.mean(scan_ms))
    # ALGORITHMIC bytes, SURVEY.md section 8(d): per tile of T reads sharing a window list, 12 B per Euler entry of the
    # list (8 B packed event + 4 B exit / dfs_end); per read a 16-B header + 4-bit alleles over its window and 8 B of
    # results (min parsimony, multiplicity; no EPP ranges are requested); per node 208 B of accumulators written once
    # and 8 B of tree arrays read once.  T = reads_per_tile (a reported design parameter; the Euler tiles of the plan).
    # The dominant kernel (the placement kernel) is charged the event stream, the reads and the per-read results; the
    # whole step adds the per-node arrays.
    win = (reads.end.astype(np.int64) - reads.start.astype(np.int64) + 1)
    b_events = st["scanned_entries"] * 12
    b_reads = int(reads.n_reads * 16 + ((win + 1) // 2).sum())
    b_results = reads.n_reads * 8
    b_nodes = arena.n_nodes * (208 + 8)
    alg = b_events + b_reads + b_results
    alg_step = alg + b_nodes
    own = (st["scanned_entries"] * (16 * 2 + 12) + reads.n_reads * (12 + 8 + 8 + 8) + reads.rm_pos.shape[0] * 5)   # round 1's own count
    # ncu counters of the same kernel on the same workload (profiles/traffic.json, captured with
    # `ncu --set full`): DRAM bytes, executed warp instructions and shared-memory wavefronts per
    # launch.  Divided by the live kernel time they say which unit the kernel is actually bound by.
    traffic, bottleneck = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            if tj.get("nodes") == arena.n_nodes and tj.get("reads_per_gpu") == reads.n_reads and tj.get("kernel") == kernel_name(st):
                traffic = tj.get("dram_bytes_per_launch")
                sm_hz = (clocks or {}).get("sm_mhz") or 1965.0
                cyc = k_ms * 1e-3 * sm_hz * 1e6
                bottleneck = {
                    "issue_slots_frac": tj["warp_instructions_per_launch"] / (cyc * 148 * 4),
                    "smem_wavefronts_frac": tj["smem_wavefronts_per_launch"] / (cyc * 148),
                    "source": tj.get("source"),
                    "note": "fractions of the SM issue rate (1 warp-instruction/cycle/sub-partition) and of the "
                            "shared-memory pipe (1 wavefront/cycle/SM) at the live kernel time"}
        except Exception:
            pass
    achieved = alg / (k_ms / 1e3) / 1e9
    out = {
        "metric": "reads placed/s", "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic", "config": config_dict(arena, reads, args),
        "read_x_node_scores_per_s": value * arena.n_nodes,
        # degree-weighted: the raw (uncollapsed) reads the placed collapsed reads stand for (rank 0's shard x N)
        "raw_reads_per_s": float(reads.degree.sum(dtype=np.int64)) * world / (ms_per_step / 1e3),
        "touched_read_entries_per_s": st["scanned_read_entries"] * 2 * world / (ms_per_step / 1e3),
        "e2e": e2e, "e2e_full_counts": e2e_full, "gpu_launches": int(st["kernel_launches"] * args.steps),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": kernel_name(st), "kernel_ms": k_ms,
                     "algorithmic_bytes_per_launch": int(alg), "peak_source": peak_src,
                     "bytes": {"events_12B_per_tile_entry": int(b_events), "reads": b_reads, "read_results": b_results,
                               "reads_per_tile": st["reads_per_tile"], "definition": "SURVEY.md section 8(d)"},
                     "step": {"algorithmic_bytes": int(alg_step), "ms": ms_per_step,
                              "achieved": alg_step / (ms_per_step / 1e3) / 1e9, "frac": alg_step / (ms_per_step / 1e3) / 1e9 / peak,
                              "note": "the whole step (placement + per-node kernels + exchange) on the section 8(d) bytes incl. N x 216 B of per-node arrays"},
                     "frac_own_definition": own / (k_ms / 1e3) / 1e9 / peak,
                     "note": "not HBM-bound, and not meant to be: the algorithmic bytes are those of scanning every tile's Euler "
                             "list (what the problem statement asks for); the kernel instead scores each DISTINCT window-restricted "
                             "haplotype of a window once per window and corrects it per read only at the states the read's few "
                             "mutations touch (~700 of 37,000 list entries), so it moves far fewer bytes than that — measured DRAM "
                             "traffic is in `traffic` — and is bound by dependent shared-memory atomics / L2 latency (DESIGN.md section 3). "
                             "frac_own_definition is round 1's 44 B per (tile, entry) count, kept for continuity",
                     "bottleneck": bottleneck},
        "roofline_node_kernels": node_roofline(st, arena, peak),
        "clocks": clocks, "c5_rescore": c5,
        "exchange": None if world == 1 else (
            "one plan on all ranks (cell histogram all-reduced in set_reads); NCCL all-reduce of the per-(bucket, state) "
            f"accumulators inside wepp_place ({x_bytes} B per step)" if shared is not None
            else ("peer-memory merge kernel (wepp_peer_merge)" if peer is not None else "NCCL all-reduce of score[N] + counts[N][50]")),
        "exchange_ms": exch_ms,
        "parity": parity, "e2e_cold": e2e_cold,
        "setup_s": {"flatten_tree": t_arena, "pack_reads_and_build_lists": t_reads},
        "stats": {k: st[k] for k in ("n_tiles", "n_lists", "n_buckets", "reads_per_tile", "stripe_width",
                                     "list_entries_total", "scanned_entries", "ms_scan_kernel", "ms_node_kernels",
                                     "place_path", "n_states", "n_window_groups")},
    }
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_reference(arena, reads, args.cpu_seconds)
        out["cpu_baseline"]["speedup_e2e"] = (e2e or {}).get("value", 0.0) / out["cpu_baseline"]["value"] if e2e else None
    print(json.dumps(out), flush=True)


def kernel_name(st) -> str:
    return {2: "delta_place_kernel", 1: "state_place_kernel", 0: "place_kernel"}[int(st["place_path"])]


def node_roofline(st, arena, peak):
    """The node-side kernel of a step (node_tile_kernel: per-node score, read counts and divergence bin count in one
    pass over node tiles) is HBM streaming.  Algorithmic bytes, SURVEY.md section 8(d): N x 208 B of per-node
    accumulators written once + N x 8 B of tree arrays (here: the 12 B per list entry of idx / enclosing boundary /
    state that stand in for them are what is actually read, counted in `read_bytes_actual`)."""
    alg = arena.n_nodes * (208 + 8)
    ms = st["ms_node_kernels"]
    ach = alg / (ms / 1e3) / 1e9 if ms > 0 else 0.0
    return {"bound": "hbm", "kernels": "node_tile_kernel (+ the 16-B accumulator packing pass)", "ms": ms,
            "algorithmic_bytes": int(alg), "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
            "read_bytes_actual": int(st["list_entries_total"] * 12),
            "note": "the counts matrix is written exactly once; the kernel is bound by the dependent look-ups "
                    "(entry -> state -> accumulator) that feed the difference tile, not by the 1.6 GB store"}


def parity_block(p, arena, reads, dev):
    """Untimed cross-checks of the step's results at the bench shape (rank 0's shard), three independent algorithms:
    the default path (sparse corrections over the distinct states), the dense state kernel and the Euler-list scan on
    the WHOLE shard — per-read integers and the full counts matrix compared on the device — and the CPU oracle on a
    sample of reads against all nodes (per-read integers of the sample; per-node score / counts of the sample placed
    alone)."""
    import torch
    import oracle
    from wepp_b200 import synth
    n, r = arena.n_nodes, reads.n_reads
    out = {"shape": f"{r} reads x {n} nodes (rank 0's shard)"}
    saved = {k: os.environ.pop(k, None) for k in ("WEPP_DELTA_PLACE", "WEPP_STATE_PLACE", "WEPP_NODE_TILES")}

    def snapshot():
        sp, sb = p.device_buffer(1)
        cp, cb = p.device_buffer(2)
        mp, mu = p.read_results()
        return (mp.copy(), mu.copy(), cuda_view(sp, sb // 8, "<f8", dev).clone(), cuda_view(cp, cb // 4, "<i4", dev).clone(),
                int(p.stats()["place_path"]))

    try:
        p.set_reads(reads)
        p.place(0, 0)
        ref = snapshot()
        out["default_path"] = ref[4]
        for name, env in (("dense_states", {"WEPP_DELTA_PLACE": "0"}), ("euler_scan", {"WEPP_STATE_PLACE": "0"}),
                          ("default_with_hbm_difference_arrays", {"WEPP_NODE_TILES": "0"})):
            for k in ("WEPP_DELTA_PLACE", "WEPP_STATE_PLACE", "WEPP_NODE_TILES"):
                os.environ.pop(k, None)
            os.environ.update(env)
            p.place(0, 0)
            cur = snapshot()
            rel = torch.max(torch.abs(cur[2] - ref[2]) / torch.clamp(torch.abs(ref[2]), min=1e-300) * (ref[2] != 0)).item()
            out[name] = {"path": cur[4], "max_parsimony_equal": bool(np.array_equal(cur[0], ref[0])),
                         "multiplicity_equal": bool(np.array_equal(cur[1], ref[1])),
                         "counts_equal": bool(torch.equal(cur[3], ref[3])), "score_max_rel_diff": float(rel),
                         "score_zero_pattern_equal": bool(torch.equal(cur[2] == 0, ref[2] == 0))}
            del cur
        for k in ("WEPP_DELTA_PLACE", "WEPP_STATE_PLACE", "WEPP_NODE_TILES"):
            os.environ.pop(k, None)
        # the CPU oracle (with the reference's range trees) on a sample spread over the shard
        n_s = 512 if n > 2_000_000 else 4096
        idx = np.unique(np.linspace(0, r - 1, n_s).astype(np.int64))
        sample = reads.take(idx)
        cores = os.cpu_count() or 1
        want_node = n * 208 * min(cores, 8) < 24e9
        t0 = time.perf_counter()
        o = oracle.cartesian_map(arena, sample, None, n_threads=min(cores, 8) if want_node else cores, want_node=want_node,
                                 range_trees=True, range_reads=reads)
        orc = {"reads": int(idx.size), "seconds": time.perf_counter() - t0,
               "max_parsimony_equal": bool(np.array_equal(o["max_parsimony"], ref[0][idx])),
               "multiplicity_equal": bool(np.array_equal(o["multiplicity"], ref[1][idx]))}
        if want_node:   # the sample placed alone: per-node results against the oracle's
            p.set_reads(sample)
            p.place(0, 0)
            sc, ct = p.node_results()
            orc["sample_path"] = int(p.stats()["place_path"])
            orc["counts_equal"] = bool(np.array_equal(ct, o["counts"]))
            nzm = o["score"] != 0
            orc["score_max_rel_diff"] = float(np.max(np.abs(sc[nzm] - o["score"][nzm]) / o["score"][nzm])) if nzm.any() else 0.0
            orc["score_zero_pattern_equal"] = bool(np.array_equal(sc == 0, o["score"] == 0))
            p.set_reads(reads)
        out["oracle_sample"] = orc
        # the distinct states are told apart by two 64-bit hashes + size: audit walk on a fresh handle (WEPP_STATE_VERIFY
        # makes wepp_place fail if any evaluated list entry differs from the entries stored for its state)
        from wepp_b200.placement import Placer
        os.environ["WEPP_STATE_VERIFY"] = "1"
        try:
            q = Placer(dev)
            q.set_arena(arena)
            q.set_reads(reads)
            q.place(0, 0)
            out["states_verified_entry_by_entry"] = {"states": int(q.stats()["n_states"]), "equal": True}
            q.close()
        except Exception as e:
            out["states_verified_entry_by_entry"] = {"equal": False, "error": repr(e)[:200]}
        finally:
            os.environ.pop("WEPP_STATE_VERIFY", None)
        flags = [v for k, d in out.items() if isinstance(d, dict) for kk, v in d.items() if kk.endswith("_equal") or kk == "equal"]
        rels = [d["score_max_rel_diff"] for d in out.values() if isinstance(d, dict) and "score_max_rel_diff" in d]
        out["all_green"] = bool(all(flags) and all(x <= 1e-9 for x in rels))
    finally:
        for k, v in saved.items():
            if v is not None:
                os.environ[k] = v
    return out


def multi_rank_checks(p, arena, reads, dev, world):
    """N > 1 (all ranks call this): the merged per-node results are the same on every rank, and read support is
    conserved over the whole job — for every count bin, sum over nodes of mapped_read_counts[v][b] equals the sum over
    ALL ranks' reads in the bin of degree x multiplicity; the score mass equals the sum of degree / (1 + parsimony).
    (The oracle cross-checks run at N = 1 and in tests/peer_worker.py on 2-4 GPUs.)"""
    import torch
    import torch.distributed as dist
    p.place(0, 0, sync=True)
    mp, mu = p.read_results()
    sp, sb = p.device_buffer(1)
    cp, cb = p.device_buffer(2)
    score = cuda_view(sp, sb // 8, "<f8", dev)
    counts = cuda_view(cp, cb // 4, "<i4", dev).view(-1, 50)
    colsum = counts.sum(dim=0, dtype=torch.int64)
    bins = np.minimum(reads.start // (arena.genome_size // 50), 49)
    expect = torch.from_numpy(np.bincount(bins, weights=reads.degree.astype(np.float64) * mu, minlength=50).astype(np.int64)).to(colsum.device)
    mass = torch.tensor([float((reads.degree / (1.0 + mp))[mu > 0].sum())], dtype=torch.float64, device=colsum.device)
    dist.all_reduce(expect)
    dist.all_reduce(mass)
    sig = torch.cat([colsum.double(), score.sum().reshape(1), (score * torch.arange(1, score.numel() + 1, device=score.device)).sum().reshape(1)])
    sigs = [torch.empty_like(sig) for _ in range(world)]
    dist.all_gather(sigs, sig)
    same = all(bool(torch.equal(s, sigs[0])) for s in sigs)
    return {"ranks": world, "merged_results_identical_on_all_ranks": same,
            "read_support_conserved": bool(torch.equal(colsum, expect)),
            "score_mass_rel_err": float(abs(score.sum().item() - mass.item()) / mass.item()),
            "all_green": bool(same and torch.equal(colsum, expect) and abs(score.sum().item() - mass.item()) <= 1e-9 * mass.item())}


def cold_sample(arena, reads, dev, q_env, k_env):
    """One sample end to end on a fresh handle, as the reference's process does it once: flattened tree to the device,
    reads in, every per-tree / per-read-set index built (Euler stripes, rank table, window lists, distinct states,
    posting lists, window groups, tile tables), placement, results out to host memory."""
    import torch
    from wepp_b200.placement import Placer
    torch.cuda.synchronize()
    ph = {}
    t_all = time.perf_counter()
    q = Placer(dev, stripe_width=q_env, reads_per_lane=k_env)
    t0 = time.perf_counter()
    q.set_arena(arena)
    ph["set_arena_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    q.set_reads(reads)
    ph["set_reads_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    q.place(0, 0)
    ph["first_place_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    q.read_results()
    q.node_summary()
    ph["results_out_s"] = time.perf_counter() - t0
    total = time.perf_counter() - t_all
    q.close()
    return {"seconds": total, "reads_per_s": reads.n_reads / total, "phases": ph,
            "note": "fresh handle: host flatten of the tree into Euler stripes + upload (set_arena), reads up + keying + window "
                    "lists (set_reads), states / posting lists / window groups / tile tables + the step itself (first_place), "
                    "per-read results and per-node score + dist_divergence to host"}


def run_c5(args, p, arena, reads, world, dev, barrier):
    """BASELINE.json configs[4]: iterative re-scoring of the rank's resident reads against a pool of 5,000
    candidate haplotypes over 10 iterations with 10 % pool churn (haplotype::mutation_distance + argmin,
    haplotype.hpp:123-177, arena.cpp:614-625).  One iteration = one wepp_rescore through the C ABI with host
    buffers: candidate indices in, per-read min distance out (and the argmin counts kept on the device);
    wall time, max over ranks."""
    import torch
    import torch.distributed as dist
    rng = np.random.default_rng(20260105)
    n_cand, iters = max(8, int(5000 * min(1.0, args.scale * 4))), 10
    pool = rng.choice(arena.n_nodes, n_cand, replace=False).astype(np.int32)
    p.set_reads(reads)
    p.rescore(pool, want_argmin=False)   # warm-up: buffers, candidate entry lists
    barrier()
    times = []
    for _ in range(iters):
        churn = rng.choice(n_cand, n_cand // 10, replace=False)
        pool[churn] = rng.integers(0, arena.n_nodes, churn.size)
        t0 = time.perf_counter()
        md, _, _, _ = p.rescore(pool, want_argmin=False)
        times.append(time.perf_counter() - t0)
    dt = float(np.sum(times))
    if world > 1:
        t = torch.tensor([dt], device=f"cuda:{dev}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    out = {"workload": f"{reads.n_reads} resident reads per GPU x {n_cand} candidate haplotypes x {iters} iterations, "
                       f"10% pool churn per iteration", "ms_per_iteration": dt / iters * 1e3,
           "read_x_candidate_distances_per_s": reads.n_reads * world * n_cand * iters / dt,
           "checksum_min_dist": int(md.astype(np.int64).sum()),
           "outputs": "min distance per read to host; argmin counts on the device"}
    if args.c5_generic:
        t0 = time.perf_counter()
        gmd, _, _, _ = p.rescore_reads(reads, pool, want_argmin=False)
        out["generic_kernel_ms_per_iteration"] = (time.perf_counter() - t0) * 1e3
        out["generic_agrees"] = bool(np.array_equal(gmd, md))
    return out


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":   # keeps NCCL's banner off stdout: one JSON line only
            del os.environ["NCCL_DEBUG"]
        dist.init_process_group("nccl", device_id=None)
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
