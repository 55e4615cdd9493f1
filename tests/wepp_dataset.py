"""A complete on-disk WEPP workspace for the CLI parity tests (TEST INFRASTRUCTURE).

Lays out what `wepp detectPeaks` reads (src/WEPP/dataset.hpp:58-113,213-220):
  <root>/data/<DIR>/<TREE>.pb[.gz]          Parsimony::data, written by the real protobuf runtime
  <root>/data/<DIR>/<REF>.fa                reference FASTA
  <root>/data/<DIR>/mask.bed                optional
  <root>/intermediate/<DIR>/<P>_reads.pb    Sam::sam, written by the real protobuf runtime
  <root>/intermediate/<DIR>/<P>_corrected_variants.tsv, <P>_depth.tsv   (only handed on to freyja)
  <root>/results/<DIR>/
  <root>/weppdir/src/Freyja/                the directory the post filter `cd`s into (post_filter.cpp:71-80)
  <root>/weppdir/src/WEPP/sam_generation.py stub of the script dump_read2haplotype_mapping runs (arena.cpp:692-695)
  <root>/bin/freyja                         a deterministic stand-in for `freyja demix` (Freyja itself — cvxpy,
                                            pandas solvers — is outside SURVEY §8 and not in this image)
The stand-in chooses haplotypes and abundances from a hash of the barcode row names and writes
freyja_output_latest.txt / residual_mutations.txt in Freyja's formats (src/Freyja/freyja/_cli.py:115-124,
sample_deconv.py:234-248), so both binaries see byte-identical deconvolution results for identical barcodes.
"""
from __future__ import annotations

import gzip
import os
import stat

import numpy as np

from oracle import formats
from wepp_b200 import synth

NUC_CHAR = {1: "A", 2: "C", 4: "G", 8: "T", 15: "N"}
NUC_IDX = {1: 0, 2: 1, 4: 2, 8: 3}

FAKE_FREYJA = r'''#!/usr/bin/env python3
"""Deterministic stand-in for `freyja demix` (tests only)."""
import hashlib, os, sys

def h(s):
    return int(hashlib.md5(s.encode()).hexdigest()[:8], 16)

def main():
    a = sys.argv[1:]
    assert a[0] == "demix", a
    variants, depth = a[1], a[2]
    opt = {a[i]: a[i + 1] for i in range(3, len(a) - 1, 2)}
    with open(os.path.join(os.path.dirname(opt["--output"]), "freyja_calls.log"), "a") as log:
        log.write(" ".join(a) + "\n")
    rows = []
    with open(opt["--barcodes"]) as f:
        header = f.readline().rstrip("\n").split(",")[1:]
        for line in f:
            rows.append(line.split(",", 1)[0])
    keep_mod = int(os.environ.get("FAKE_FREYJA_KEEP_MOD", "3"))
    chosen = [r for r in rows if h(r) % keep_mod == 0] or rows[:1]
    ab = [1 + h(r + "x") % 1000 for r in chosen]
    order = sorted(range(len(chosen)), key=lambda i: (-ab[i], chosen[i]))
    chosen = [chosen[i] for i in order]
    tot = float(sum(ab))
    ab = [ab[i] / tot * 0.97 for i in order]
    with open(opt["--output"], "w") as f:
        f.write("\tsample\n")
        f.write("summarized\t[('Other', 0.97)]\n")
        f.write("lineages\t" + " ".join(chosen) + "\n")
        f.write("abundances\t" + " ".join("%.8f" % x for x in ab) + "\n")
        f.write("resid\t1.0\n")
        f.write("coverage\t99.0\n")
    # residual mutations: a few barcode columns, as "<pos><alt>" (allele under-explained) or
    # "<pos><ref>" (reference allele under-explained)
    out_dir = os.path.dirname(opt["--output"])
    n_res = int(os.environ.get("FAKE_FREYJA_RESIDUALS", "12"))
    picked = sorted(header, key=lambda m: h(m))[:n_res]
    with open(os.path.join(out_dir, "residual_mutations.txt"), "w") as f:
        for k, m in enumerate(picked):
            if not m or m[-1] not in "ACGT":
                continue
            mut = (m[1:-1] + m[0]) if k % 3 == 0 else m[1:]
            f.write("%s,%.6f,%.4f,%d\n" % (mut, 0.01 * (k + 1), 0.05 * (k % 7 + 1), 100 + k))

main()
'''

SAM_GENERATION_STUB = '''import sys
open(sys.argv[1] + "/" + sys.argv[3] + "_sam_generation_called.txt", "w").write(" ".join(sys.argv[1:]) + "\\n")
'''


def _newick(parent, ids):
    n = len(parent)
    kids = [[] for _ in range(n)]
    for v in range(1, n):
        kids[parent[v]].append(v)
    out = []
    # iterative preorder emission
    stack = [(0, 0)]
    while stack:
        v, state = stack.pop()
        if state == 0:
            if kids[v]:
                out.append("(")
                stack.append((v, 1))
                for i, c in enumerate(reversed(kids[v])):
                    stack.append((c, 0))
                    if i != len(kids[v]) - 1:
                        stack.append((-1, 2))
            else:
                out.append(ids[v] + ":1")
        elif state == 1:
            out.append(")" + (":1" if v else ""))
        else:
            out.append(",")
    return "".join(out) + ";"


def make_workspace(root, *, n_nodes=2500, genome=3000, n_reads=3000, seed=5, dataset="ds", prefix="smp",
                   tree_name="tree.pb.gz", ref_name="ref.fa", with_mask=True, n_amplicons=12, read_len=150,
                   n_templates=40, clade_levels=2):
    """Returns a dict of paths and the generated arrays."""
    root = str(root)
    arena = synth.make_arena(n_nodes, genome, seed, mean_depth=12.0, root_events=0)
    reads = synth.make_reads(arena, n_reads, seed, amplicons=synth.amplicon_scheme(genome, n_amplicons, 300, 400, seed),
                             read_len=read_len, n_templates=n_templates)
    rng = np.random.default_rng(seed + 99)
    n = arena.n_nodes
    parent = arena.parent.tolist()
    is_leaf = np.ones(n, bool)
    is_leaf[arena.parent[1:]] = False
    ids = [f"S{v}|hap/{v}" if is_leaf[v] else f"node_{v}" for v in range(n)]   # internal ids are renamed by the loader
    ref = "".join(NUC_CHAR[int(c)] for c in arena.ref_codes[1:])

    data = formats.ParsimonyData()
    data.newick = _newick(parent, ids)
    # par_nuc: the allele on the root path before this event
    cur_allele = [dict() for _ in range(n)]     # small trees only
    for v in range(n):
        ml = data.node_mutations.add()
        inherited = dict(cur_allele[parent[v]]) if v else {}
        for k in range(int(arena.mut_off[v]), int(arena.mut_off[v + 1])):
            pos, nuc, rf = int(arena.mut_pos[k]), int(arena.mut_nuc[k]), int(arena.mut_ref[k])
            par = inherited.get(pos, rf)
            m = ml.mutation.add()
            m.position = pos
            m.ref_nuc = NUC_IDX[rf]
            m.par_nuc = NUC_IDX[par] if par in NUC_IDX else NUC_IDX[rf]
            m.mut_nuc.extend([i for i in range(4) if nuc & (1 << i)])
            m.chromosome = "chr"
            inherited[pos] = nuc
        cur_allele[v] = inherited
        md = data.metadata.add()
        ann = []
        for lvl in range(clade_levels):
            p_named = 0.02 if v else 1.0
            ann.append(f"L{lvl}.{int(rng.integers(50))}" if rng.random() < p_named else "")
        md.clade_annotations.extend(ann)
    pb = data.SerializeToString()

    ddir = os.path.join(root, "data", dataset)
    idir = os.path.join(root, "intermediate", dataset)
    rdir = os.path.join(root, "results", dataset)
    for d in (ddir, idir, rdir, os.path.join(root, "weppdir", "src", "Freyja"), os.path.join(root, "weppdir", "src", "WEPP"),
              os.path.join(root, "bin")):
        os.makedirs(d, exist_ok=True)
    tree_path = os.path.join(ddir, tree_name)
    with (gzip.open(tree_path, "wb") if ".gz" in tree_name else open(tree_path, "wb")) as f:
        f.write(pb)
    with open(os.path.join(ddir, ref_name), "w") as f:
        f.write(">chrREF some description\n")
        for i in range(0, len(ref), 70):
            line = ref[i:i + 70]
            f.write((line.lower() if (i // 70) % 5 == 4 else line) + "\n")
    masked = []
    if with_mask:
        masked = sorted(rng.choice(np.arange(1, genome + 1), size=max(genome // 100, 1), replace=False).tolist())
        with open(os.path.join(ddir, "mask.bed"), "w") as f:
            for p in masked:
                f.write(f"chrREF\t{p - 1}\t{p}\n")

    sam = formats.SamSam()
    for i in range(reads.n_reads):
        s, e = int(reads.start[i]), int(reads.end[i])
        content = list(ref[s - 1:e])
        for k in range(int(reads.rm_off[i]), int(reads.rm_off[i + 1])):
            content[int(reads.rm_pos[k]) - s] = NUC_CHAR[int(reads.rm_nuc[k])]
        if i % 17 == 3 and len(content) > 20:     # a paired-read gap
            for k in range(8, 14):
                content[k] = "_"
        r = sam.reads.add()
        deg = int(reads.degree[i])
        r.read = f"q{i}_READ_{s}_{e}_{deg}"
        r.start_idx = s
        r.content = "".join(content)
        r.degree = deg
        col = sam.reverse_columns.add()
        col.column_name = r.read
        col.input_columns.extend([f"raw{i}.{j}" for j in range(deg)])
    with open(os.path.join(idir, f"{prefix}_reads.pb"), "wb") as f:
        f.write(sam.SerializeToString())
    for name in (f"{prefix}_corrected_variants.tsv", f"{prefix}_depth.tsv"):
        with open(os.path.join(idir, name), "w") as f:
            f.write("placeholder\n")
    fr = os.path.join(root, "bin", "freyja")
    with open(fr, "w") as f:
        f.write(FAKE_FREYJA)
    os.chmod(fr, os.stat(fr).st_mode | stat.S_IXUSR | stat.S_IXGRP | stat.S_IXOTH)
    with open(os.path.join(root, "weppdir", "src", "WEPP", "sam_generation.py"), "w") as f:
        f.write(SAM_GENERATION_STUB)
    return {"root": root, "dataset": dataset, "prefix": prefix, "tree": tree_name, "ref": ref_name,
            "wepp_dir": os.path.join(root, "weppdir"), "bin": os.path.join(root, "bin"), "idir": idir, "rdir": rdir,
            "arena": arena, "reads": reads, "masked": masked, "reference": ref}


def cli_args(ws, threads=4, clade_idx=1, min_af="0.005", min_prop="0.005"):
    return ["detectPeaks", "-w", ws["wepp_dir"], "-T", str(threads), "-i", ws["tree"], "-p", ws["prefix"], "-f", ws["ref"],
            "-d", ws["dataset"], "-a", min_af, "-r", min_prop, "-n", str(clade_idx)]


def env_with_fake_freyja(ws):
    env = dict(os.environ)
    env["PATH"] = ws["bin"] + os.pathsep + env.get("PATH", "")
    return env
