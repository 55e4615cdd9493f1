#!/usr/bin/env python
"""tests/golden/c1_full.npz: the FULL C1 configuration (BASELINE.json configs[0] shape: 50,000-node tree, genome
15,222, 200,000 collapsed reads on the RSV-A primer scheme; synth.config_shape("C1")) through the REFERENCE'S OWN
object code (oracle/_ref/libwepp_ref.so: arena + cartesian_map, initial_filter.cpp:139-239).

Run where the reference tree is mounted:   python tests/golden/make_c1.py
Stored: per-read max_parsimony / multiplicity, per-node score and dist_divergence, of mapped_read_counts[N][50] the
SHA-256 of the int32 matrix (bit-exact equality), its column sums and every 16th row (to localise a mismatch), and a
digest of the inputs (regenerated from the seed by the test)."""
from __future__ import annotations

import hashlib
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def inputs():
    from wepp_b200 import synth
    arena, reads, _ = synth.config_shape("C1")
    return arena, reads


def digest(arena, reads) -> str:
    h = hashlib.sha256()
    for a in (arena.parent, arena.mut_off, arena.mut_pos, arena.mut_ref, arena.mut_nuc, arena.ref_codes, reads.start,
              reads.end, reads.degree, reads.rm_off, reads.rm_pos, reads.rm_nuc):
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def cover_reads(arena, reads):
    """All-reference cover reads (one per 150 bases) make every site covered, so the reference's condensed arena is
    exactly the arena handed in; they come last and are not placed."""
    from wepp_b200 import synth
    g = arena.genome_size
    cs = np.arange(1, g + 1, 150, dtype=np.int32)
    ce = np.minimum(cs + 149, g).astype(np.int32)
    return synth.Reads(np.concatenate([reads.start, cs]), np.concatenate([reads.end, ce]),
                       np.concatenate([reads.degree, np.ones(cs.size, np.int32)]),
                       np.concatenate([reads.rm_off, np.full(cs.size, reads.rm_off[-1], np.int64)]),
                       reads.rm_pos, reads.rm_nuc)


if __name__ == "__main__":
    from oracle import ref as oref
    arena, reads = inputs()
    t0 = time.time()
    sess = oref.Session(arena, cover_reads(arena, reads), threads=os.cpu_count() or 1)
    assert sess.n_nodes == arena.n_nodes, (sess.n_nodes, arena.n_nodes)
    ra = sess.arena()
    assert np.array_equal(ra["parent"], arena.parent) and np.array_equal(ra["mut_pos"], arena.mut_pos)
    out = sess.cartesian_map(n_sel=reads.n_reads, want_node=True, want_epp=False)
    print(f"reference: arena + cartesian_map of {reads.n_reads} reads x {arena.n_nodes} nodes in {time.time() - t0:.1f} s "
          f"(cartesian mapping took {out['ms']:.0f} ms)")
    ct = np.ascontiguousarray(out["counts"], dtype=np.int32)
    np.savez_compressed(os.path.join(HERE, "c1_full.npz"), digest=np.array(digest(arena, reads)),
                        max_parsimony=out["max_parsimony"].astype(np.uint8), multiplicity=out["multiplicity"],
                        score=out["score"], counts_sha256=np.array(hashlib.sha256(ct.tobytes()).hexdigest()),
                        counts_colsum=ct.sum(axis=0, dtype=np.int64), counts_rows_every_16=ct[::16].copy(),
                        dist_divergence=out["dist_divergence"])
    print("max parsimony", int(out["max_parsimony"].max()), "file", os.path.getsize(os.path.join(HERE, "c1_full.npz")) / 1e6, "MB")
