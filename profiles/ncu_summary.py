#!/usr/bin/env python
"""Text summary of an ncu report for profiles/: per kernel the headline raw metrics, the stall-reason split and the
hottest SASS lines.  usage: python profiles/ncu_summary.py X.ncu-rep > profiles/NAME.txt   (ncu must be on PATH)"""
import csv, io, subprocess, sys
rep = sys.argv[1]
WANT = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_atom.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
print(f"# {rep}: ncu --set full --clock-control none --import-source on (one launch per kernel, steady state)")
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print(f"\n== {d.get('Kernel Name', '?')}  (launch id {d.get('ID', '?')})")
    for k in WANT:
        if k in d:
            print(f"  {k:78s} {d[k]:>16s} {units[hdr.index(k)]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
# the source page concatenates the kernels: "Kernel Name" rows start a section
sec, name = [], None
def flush(name, sec):
    if not sec:
        return
    h = sec[0]
    col = {x: i for i, x in enumerate(h)}
    stalls = [x for x in h if x.startswith("stall_") and "Not Issued" not in x]
    body = [r for r in sec[1:] if len(r) >= len(h)]
    tot = {s: sum(int(r[col[s]] or 0) for r in body) for s in stalls}
    ns = sum(int(r[col["# Samples"]] or 0) for r in body)
    ni = sum(int(r[col["Instructions Executed"]] or 0) for r in body)
    print(f"\n== {name}: {ns} samples, {ni} warp instructions")
    for s, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]:
        if v:
            print(f"  {s:28s} {100.0 * v / max(ns, 1):5.1f} %")
    print("  hottest SASS lines (samples, executed, top stall):")
    for r in sorted(body, key=lambda r: -int(r[col["# Samples"]] or 0))[:14]:
        st = max(stalls, key=lambda s: int(r[col[s]] or 0))
        print(f"  {int(r[col['# Samples']]):8d} {int(r[col['Instructions Executed']]):11d}  {st:22s} {r[col['Source']].strip()[:80]}")
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Kernel Name":
        flush(name, sec)
        name, sec = r[1] if len(r) > 1 else "?", []
    else:
        sec.append(r)
flush(name, sec)
