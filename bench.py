#!/usr/bin/env python
"""bench.py — reads placed/s of WEPP's parsimonious read placement on B200(s).

    python bench.py --gpus N --steps K --warmup W              # our CUDA path
    python bench.py --impl reference --gpus N --steps K ...    # the reference's CPU placement

Workload (BASELINE.json configs[2], the one the metric is quoted on): a synthetic
public-scale SARS-CoV-2 MAT (8M arena nodes, genome 29,903) x ARTIC-like 150-bp amplicon
reads, read-sharded: every GPU places 1.25M collapsed reads against the replicated tree
(weak scaling; 8 GPUs = the full 10M reads).  One step = one cartesian_map over the rank's
shard: placement kernel (state_place_kernel: the distinct window-restricted haplotypes of every
window scored once) + expansion + per-node scans, and for N>1 the exchange step (one kernel over
NVLink peer memory, or --exchange nccl: the all-reduce of the per-node score / read-count arrays).

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

NODES_FULL = 8_000_000
READS_PER_GPU_FULL = 1_250_000
GENOME = 29903


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink nodes and reads (development only)")
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="CPU-baseline budget per measurement")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N>1: per-node exchange step — fused peer-memory merge kernel (default) or NCCL all-reduce")
    ap.add_argument("--no-c5", action="store_true", help="skip the candidate re-scoring (C5) measurement")
    ap.add_argument("--c5-generic", action="store_true", help="also time the generic K4 kernel once (slow)")
    return ap.parse_args()


def workload(scale: float, rank: int):
    from wepp_b200 import synth
    n_nodes = max(int(NODES_FULL * scale), 1000)
    n_reads = max(int(READS_PER_GPU_FULL * scale), 256)
    arena = synth.make_arena(n_nodes, GENOME, synth.SEED)
    reads = synth.make_reads(arena, n_reads, synth.SEED + 1000 * rank)
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
REF_MAX_NODES = 2_500_000   # above this the reference's own arena build (O(N*depth) stack_muts, ~5 KB/node) and its
                            # per-chunk N x 208 B score arrays (initial_filter.cpp:155-158) do not fit a bounded sample


def cpu_reference(arena, reads, seconds: float, repeats: int = 1):
    """The reference's CPU placement on a bounded read sample with all host threads: the
    shim-compiled reference object code (oracle/_ref) when it is present and the tree is small
    enough for its arena to be built inside the time/memory budget, else the oracle port
    (measured 1.23x slower than the reference's own code at 0.7M nodes / 8 threads)."""
    import oracle
    from wepp_b200 import synth
    cores = os.cpu_count() or 1
    kind, sess, t_build = "port", None, 0.0
    try:
        from oracle import ref as oref
        if oref.available() and arena.n_nodes <= REF_MAX_NODES and oref.fits_in_memory(arena.n_nodes):
            # all-reference cover reads (one per 150 bases) make every site covered, so the
            # reference's condensed arena is exactly the arena our side places against
            g = arena.genome_size
            cs = np.arange(1, g + 1, 150, dtype=np.int32)
            ce = np.minimum(cs + 149, g).astype(np.int32)
            cover = synth.Reads(np.concatenate([reads.start, cs]), np.concatenate([reads.end, ce]),
                                np.concatenate([reads.degree, np.ones(cs.size, np.int32)]),
                                np.concatenate([reads.rm_off, np.full(cs.size, reads.rm_off[-1], np.int64)]),
                                reads.rm_pos, reads.rm_nuc)
            t0 = time.perf_counter()
            sess = oref.Session(arena, cover, threads=cores)
            t_build = time.perf_counter() - t0
            if sess.n_nodes == arena.n_nodes:
                kind = "reference"
    except Exception:
        sess = None

    def run(n):
        t0 = time.perf_counter()
        if kind == "reference":
            sess.cartesian_map(n_sel=n, want_node=False, want_epp=False)
        else:
            oracle.cartesian_map(arena, reads.slice(0, n), None, n_threads=cores, want_node=False)
        return max(time.perf_counter() - t0, 1e-6)

    # calibrate on a few reads per thread, then size the sample for `seconds`
    # (two rounds: the first small run carries one-off costs — thread start-up, page faults of the
    # per-thread node arrays — and would undersize the sample)
    n0 = min(reads.n_reads, cores * (16 if kind == "reference" else 1))
    dt = run(n0)
    n_mid = int(min(reads.n_reads, max(n0, n0 * min(seconds / 8.0, 3.0) / dt)))
    n_mid = min(max(cores, (n_mid // cores) * cores), reads.n_reads)
    if n_mid > n0:
        dt, n0 = run(n_mid), n_mid
    n1 = int(min(reads.n_reads, max(n0, n0 * seconds / dt)))
    n1 = min(max(cores, (n1 // cores) * cores), reads.n_reads)
    dts = [run(n1) for _ in range(max(1, repeats))]
    dt = float(np.median(dts))
    if dt < 0.5 * seconds and n1 < reads.n_reads:
        # still short of the budget (the per-thread N x 208 B arrays dominate small samples): one more round
        n1 = int(min(reads.n_reads, n1 * seconds / dt))
        n1 = min(max(cores, (n1 // cores) * cores), reads.n_reads)
        dts = [run(n1) for _ in range(max(1, repeats))]
        dt = float(np.median(dts))
    return {"value": n1 / dt, "unit": "reads/s", "cores": cores, "kind": kind,
            "sample": f"{n1} of the step's {reads.n_reads} reads vs all {arena.n_nodes} nodes, {dt:.1f} s"
                      + (f" (+{t_build:.1f} s reference arena build, untimed)" if kind == "reference" else "")}


def run_reference(args, rank: int):
    if rank != 0:
        return
    arena, reads = workload(args.scale, 0)
    base = cpu_reference(arena, reads, args.cpu_seconds, repeats=max(1, min(args.steps, 3)))
    v = float(base["value"])
    out = {"impl": "reference", "metric": "reads placed/s", "value": v, "unit": "reads/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
           "config": config_dict(arena, reads, args), "cpu_baseline": base,
           "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def config_dict(arena, reads, args):
    return {"workload": f"C3 synthetic SARS-CoV-2-scale MAT ({arena.n_nodes} arena nodes, {arena.n_events} events, "
                        f"genome {GENOME}) x {reads.n_reads} ARTIC-like 150-bp collapsed reads per GPU "
                        f"(read-sharded; 8 GPUs = 10M reads)",
            "nodes": arena.n_nodes, "reads_per_gpu": reads.n_reads, "seed": 20260101,
            "l2": "inputs exceed L2 (Euler lists + 1.6 GB per-node arrays are re-zeroed and re-streamed every step)"}


# ---------------------------------------------------------------------------------------------
def run_ours(args, rank: int, world: int, local_rank: int):
    import torch
    import torch.distributed as dist
    from wepp_b200 import multigpu, synth
    from wepp_b200.placement import Placer

    torch.cuda.set_device(local_rank)
    dev = local_rank
    arena, reads = workload(args.scale, rank)
    k_env = int(os.environ.get("WEPP_READS_PER_LANE", "0"))   # development knob
    q_env = int(os.environ.get("WEPP_STRIPE_WIDTH", "16"))
    p = Placer(dev, stripe_width=q_env, reads_per_lane=k_env)
    stream = torch.cuda.current_stream()
    p.set_stream(stream.cuda_stream)
    t0 = time.perf_counter()
    p.set_arena(arena)
    t_arena = time.perf_counter() - t0
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
        if world == 1:
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
    if world > 1:
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
               "note": "every step re-uploads the reads, keys and buckets them and rebuilds the per-window Euler lists; the "
                       "per-tree indices (stripe rank table, 2.7 ms; distinct window-restricted haplotypes of the window "
                       "lists, 42 ms) are built on the first call and kept while read sets map to the same windows and bins",
               "phases": e2e_phases()}
        dt = time_e2e(True)
        e2e_full = {"value": reads_total / dt, "unit": "reads/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": int(r * 8 + n * (8 + 200) + d2h_keys), "ms_per_step": dt * 1e3,
                    "outputs": "as e2e, with mapped_read_counts[N][50] instead of dist_divergence"}

    if peer is not None:
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
    # algorithmic bytes of the placement kernel alone: two passes over each tile's Euler list
    # (16-B entries), one 12-B accumulator update per (tile, entry), packed reads in, results out
    alg = (st["scanned_entries"] * (16 * 2 + 12) + reads.n_reads * (12 + 8 + 8 + 8) + reads.rm_pos.shape[0] * 5)
    # ncu counters of the same kernel on the same workload (profiles/traffic.json, captured with
    # `ncu --set full`): DRAM bytes, executed warp instructions and shared-memory wavefronts per
    # launch.  Divided by the live kernel time they say which unit the kernel is actually bound by.
    traffic, bottleneck = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            if tj.get("nodes") == arena.n_nodes and tj.get("reads_per_gpu") == reads.n_reads:
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
                     "traffic": traffic, "kernel": "state_place_kernel" if os.environ.get("WEPP_STATE_PLACE", "1") != "0" else "place_kernel",
                     "kernel_ms": k_ms,
                     "algorithmic_bytes_per_launch": int(alg), "peak_source": peak_src,
                     "note": "not HBM-bound: the per-window lists are L2-resident and every tile re-reads them; the kernel "
                             "(state_place_kernel: the distinct window-restricted haplotypes of a window are scored once "
                             "each, about half the entries of the Euler list the algorithmic bytes are counted on) is "
                             "limited by warp-instruction issue (see bottleneck and DESIGN.md section 3)",
                     "bottleneck": bottleneck},
        "roofline_node_kernels": node_roofline(st, arena, peak),
        "clocks": clocks, "c5_rescore": c5,
        "exchange": None if world == 1 else ("peer-memory merge kernel (wepp_peer_merge)" if peer is not None
                                             else "NCCL all-reduce of score[N] + counts[N][50]"),
        "exchange_ms": exch_ms,
        "setup_s": {"flatten_tree": t_arena, "pack_reads_and_build_lists": t_reads},
        "stats": {k: st[k] for k in ("n_tiles", "n_lists", "n_buckets", "reads_per_tile", "stripe_width",
                                     "list_entries_total", "scanned_entries", "ms_scan_kernel", "ms_node_kernels")},
    }
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_reference(arena, reads, args.cpu_seconds)
    print(json.dumps(out), flush=True)


def node_roofline(st, arena, peak):
    """The node-side kernels of a step (expand_kernel + the score / count scans, 6-7 % of it) are plain HBM
    streaming.  Algorithmic bytes: the per-node accumulators written once (N x (8 + 200) B), the per-bucket entry
    accumulators (12 B), list entries (16 B) and enclosing-boundary indices (4 B) read once."""
    acc_total = st["list_entries_total"] * st["n_buckets"] / max(st["n_lists"], 1)
    alg = arena.n_nodes * 208 + acc_total * 12 + st["list_entries_total"] * 20
    ms = st["ms_node_kernels"]
    ach = alg / (ms / 1e3) / 1e9 if ms > 0 else 0.0
    return {"bound": "hbm", "kernels": "expand_kernel + score/count difference-array scans", "ms": ms,
            "algorithmic_bytes": int(alg), "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
            "note": "the difference-array formulation moves ~4x the algorithmic bytes (zero, scatter, sum pass, apply pass)"}


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
