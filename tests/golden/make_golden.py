#!/usr/bin/env python
"""Generate tests/golden/*.npz from the REFERENCE'S OWN object code (oracle/_ref/libwepp_ref.so,
shim-compiled from /root/reference/src/WEPP by oracle/Makefile).

Run where the reference tree is mounted:   python tests/golden/make_golden.py
Every case is regenerated from a seed by tests/cases.py (inputs are NOT stored, only a digest),
run through the reference in a fresh subprocess (the reference allows one data set per process)
and its outputs are stored: the flattened arena (arena.cpp:3-56), the masked reads, the
cartesian_map results (initial_filter.cpp:139-239), single_read_tree under a mapped mask
(:112-135), haplotype::mutation_distance (haplotype.hpp:123-177) and wepp_filter::filter (:455-506).
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# name -> (kind, seed, kwargs)
CASES = {
    "tiny0": ("tiny", 0, {}),
    "tiny1": ("tiny", 1, {}),
    "tiny3": ("tiny", 3, {}),
    "tiny5_dense": ("tiny", 5, {"n_nodes": 60, "genome": 64, "n_reads": 90, "max_len": 50}),
    "small7": ("small", 7, {}),
    "small11_long": ("small", 11, {"n_nodes": 2000, "n_reads": 300, "genome": 4000}),
}


def make_inputs(name: str):
    from tests import cases
    kind, seed, kw = CASES[name]
    if kind == "tiny":
        tree, reads, mapped = cases.tiny_case(seed, **kw)
    else:
        tree, reads = cases.small_case(seed=seed, **kw)
        rng = np.random.default_rng(seed + 99)
        mapped = (rng.random(tree.n_nodes) < 0.2).astype(np.uint8)
    rng = np.random.default_rng(seed + 7)
    masked = np.sort(rng.choice(np.arange(1, tree.genome_size + 1), size=max(2, tree.genome_size // 40),
                                replace=False)).astype(np.int32)
    return tree, reads, masked, mapped


def digest(tree, reads, masked) -> str:
    h = hashlib.sha256()
    for a in (tree.parent, tree.mut_off, tree.mut_pos, tree.mut_ref, tree.mut_nuc, tree.ref_codes, reads.start,
              reads.end, reads.degree, reads.rm_off, reads.rm_pos, reads.rm_nuc, masked):
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def run_reference(name: str):
    from oracle import ref
    tree, reads, masked, mapped_mat = make_inputs(name)
    s = ref.Session(tree, reads, masked=masked, threads=4)
    out = {"digest": np.array(digest(tree, reads, masked))}
    a = s.arena()
    for k, v in a.items():
        out["arena_" + k] = v
    ro, rp, rn = s.masked_reads()
    out.update(reads_rm_off=ro, reads_rm_pos=rp, reads_rm_nuc=rn)
    r = s.cartesian_map()
    for k in ("max_parsimony", "multiplicity", "score", "counts", "dist_divergence", "epp_off", "epp_nodes"):
        out["cm_" + k] = r[k]
    # single_read_tree under a mapped mask (mask defined on arena nodes through their source MAT node)
    mapped = mapped_mat[a["source"]].astype(np.uint8)
    out["srt_mapped"] = mapped
    sel = np.unique(np.linspace(0, reads.n_reads - 1, 12).astype(np.int64))
    vals, offs, nodes = [], [0], []
    for ri in sel:
        mv, nd = s.single_read_tree(int(ri), mapped)
        vals.append(mv)
        nodes.append(nd)
        offs.append(offs[-1] + nd.size)
    out.update(srt_reads=sel, srt_max_val=np.array(vals, np.int32), srt_off=np.array(offs, np.int64),
               srt_nodes=np.concatenate(nodes).astype(np.int32) if nodes else np.zeros(0, np.int32))
    rng = np.random.default_rng(5)
    cand = np.sort(rng.choice(s.n_nodes, size=min(s.n_nodes, 64), replace=False)).astype(np.int32)
    out["md_cand"] = cand
    out["md_dist"] = s.mutation_distance(cand)
    out["filter_selected"] = s.filter()
    return out


def main():
    from oracle import ref
    if not ref.available():
        raise SystemExit("oracle/_ref/libwepp_ref.so missing: run `make -C oracle ref` where /root/reference is mounted")
    for name in CASES:
        out = ref.run_case("tests.golden.make_golden", "run_reference", name)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: arena {out['arena_parent'].shape[0]} nodes, {out['cm_max_parsimony'].shape[0]} reads, "
              f"{os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
