#!/usr/bin/env python
"""Whole-stage drop-in timing: `build/wepp detectPeaks` (this repository, GPU) against the reference's own binary
(oracle/_ref/wepp_ref: the reference's translation units, shim-compiled — see oracle/Makefile) on the same on-disk
workspace, same command line, all host threads; every output file is compared afterwards (tests/test_cli_gpu.py's
comparison).  Freyja is an external tool on both sides and is replaced by the deterministic stand-in of
tests/wepp_dataset.py.   usage: python profiles/cli_stage_timing.py [n_nodes] [n_reads]   -> one JSON line"""
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import wepp_dataset as wd          # noqa: E402
from tests.test_cli_gpu import OURS, REF, compare_workspaces   # noqa: E402


def run(binary, ws, root, threads):
    env = wd.env_with_fake_freyja(dict(ws, bin=os.path.join(root, "bin")))
    args = wd.cli_args(dict(ws, wepp_dir=os.path.join(root, "weppdir")), threads=threads)
    t0 = time.perf_counter()
    r = subprocess.run([binary] + args, cwd=root, env=env, capture_output=True, text=True, timeout=3000)
    dt = time.perf_counter() - t0
    if r.returncode != 0:
        sys.stderr.write(r.stdout[-3000:] + r.stderr[-3000:])
        raise SystemExit(f"{binary} failed")
    stages = [l for l in r.stdout.splitlines() if "took" in l.lower() or "elapsed" in l.lower()]
    stages += [l for l in r.stderr.splitlines() if l.startswith("[wepp timing]")]
    return dt, stages


def main():
    n_nodes = int(sys.argv[1]) if len(sys.argv) > 1 else 150_000
    n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 40_000
    only = os.environ.get("ONLY", "")
    threads = os.cpu_count() or 1
    tmp = tempfile.mkdtemp(prefix="wepp_cli_")
    a, b = os.path.join(tmp, "ref"), os.path.join(tmp, "ours")
    t0 = time.perf_counter()
    ws = wd.make_workspace(a, n_nodes=n_nodes, genome=29903, n_reads=n_reads, seed=17, n_templates=60, n_amplicons=99)
    shutil.copytree(a, b)
    t_make = time.perf_counter() - t0
    out = {"workload": f"detectPeaks on a {n_nodes}-node MAT (genome 29903) x {n_reads} collapsed 150-bp reads, "
                       f"99 amplicons; -T {threads}", "host_threads": threads, "make_workspace_s": round(t_make, 1)}
    c = os.path.join(tmp, "warm")
    shutil.copytree(a, c)
    if only != "ours":
        out["reference_binary_s"], out["reference_stage_lines"] = run(REF, ws, a, threads)
    if only != "ref":
        run(OURS, ws, c, threads)                       # warm-up copy: pays the page-in of the CUDA libraries once
        out["ours_s"], out["ours_stage_lines"] = run(OURS, ws, b, threads)
        if only != "ours":
            compare_workspaces(a, b, ws)
            out["outputs_identical"] = True
            out["speedup"] = out["reference_binary_s"] / out["ours_s"]
    print(json.dumps(out), flush=True)
    shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
