"""`build/wepp sam2PB` against the reference's own sam2PB object code
(oracle/_ref/wepp_ref, shim-compiled from src/WEPP/sam2pb.cpp) on synthetic SAM files: CIGAR walk with
M/I/D/S/H/N operations, Phred masking, depth and allele-frequency masking, duplicate collapsing, the
reverse-merge table.  Both outputs are decoded by the real protobuf runtime (oracle/formats.py).

The files compared with the reference hold no header / unmapped lines: the reference leaves its TBB
sub-range at the first such line (src/WEPP/sam2pb.cpp:165-168) BEFORE the reads parsed so far in that
sub-range are appended (:259-278), so it loses a chunking-dependent number of reads around every such line
(with `samtools view -h` input, workflow/rules/sam2pb.smk:9: the reads sharing a sub-range with the header).
The product skips such lines wherever they are and loses nothing (test_header_lines_anywhere).  Which raw read a
collapsed read is named after depends on thread completion order in the reference (unstable sort of equal
reads), so names are compared as "one of the group" and reverse-merge groups as sets.  No GPU needed."""
import os
import subprocess

import numpy as np
import pytest

from oracle import formats
from wepp_b200 import io as wio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OURS = os.path.join(ROOT, "build", "wepp")
REF = os.path.join(ROOT, "oracle", "_ref", "wepp_ref")
needs_ref = pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/wepp_ref not built (no /root/reference here)")
pytestmark = pytest.mark.skipif(not os.path.exists(OURS), reason="build/wepp not built")


def make_sam_workspace(root, seed, n_templates=60, copies=6, genome=1200, tail_junk=True, head_junk=False):
    rng = np.random.default_rng(seed)
    ref = "".join(rng.choice(list("ACGT"), size=genome))
    ddir, idir = os.path.join(root, "data", "ds"), os.path.join(root, "intermediate", "ds")
    os.makedirs(ddir, exist_ok=True)
    os.makedirs(idir, exist_ok=True)
    with open(os.path.join(ddir, "ref.fa"), "w") as f:
        f.write(">refseq test\n")
        for i in range(0, genome, 60):
            f.write(ref[i:i + 60] + "\n")
    lines = []
    qn = 0
    for t in range(n_templates):
        start = int(rng.integers(1, genome - 260))
        # a CIGAR made of random operations; the aligned part stays inside the genome
        ops = []
        if rng.random() < 0.3:
            ops.append((int(rng.integers(1, 8)), "S"))
        if rng.random() < 0.1:
            ops.insert(0, (int(rng.integers(1, 5)), "H"))
        ref_used = 0
        while ref_used < 120:
            ops.append((int(rng.integers(20, 60)), "M"))
            ref_used += ops[-1][0]
            u = rng.random()
            if u < 0.25:
                ops.append((int(rng.integers(1, 4)), "I"))
            elif u < 0.5:
                ops.append((int(rng.integers(1, 6)), "D"))
                ref_used += ops[-1][0]
            elif u < 0.55:
                ops.append((int(rng.integers(2, 5)), "N"))
        if ops[-1][1] != "M":
            ops.append((int(rng.integers(5, 20)), "M"))
        if rng.random() < 0.3:
            ops.append((int(rng.integers(1, 8)), "S"))
        seq_len = sum(n for n, o in ops if o in "MIS") + sum(n for n, o in ops if o == "N")
        # sequence follows the reference over M (so that duplicates and majority alleles exist), with SNPs
        seq = []
        pos = start - 1
        for n, o in ops:
            if o == "M":
                seq.extend(ref[pos:pos + n])
                pos += n
            elif o == "D":
                pos += n
            elif o in "ISN":
                seq.extend(rng.choice(list("ACGT"), size=n))
        seq = seq[:seq_len] + list(rng.choice(list("ACGT"), size=max(0, seq_len - len(seq))))
        for _ in range(int(rng.integers(0, 3))):
            seq[int(rng.integers(len(seq)))] = str(rng.choice(list("ACGT")))
        cigar = "".join(f"{n}{o}" for n, o in ops)
        for c in range(copies + int(rng.integers(0, 8))):
            s = list(seq)
            q = ["I"] * len(s)
            if rng.random() < 0.4:                     # a low-quality base or an error: breaks the duplicate group
                k = int(rng.integers(len(s)))
                if rng.random() < 0.5:
                    q[k] = chr(33 + int(rng.integers(0, 20)))
                else:
                    s[k] = str(rng.choice(list("ACGTRn")))
            flag = 0 if rng.random() < 0.5 else 16
            lines.append(f"read{qn}\t{flag}\trefseq\t{start}\t60\t{cigar}\t*\t0\t0\t{''.join(s)}\t{''.join(q)}\tNM:i:1")
            qn += 1
    order = rng.permutation(len(lines))
    lines = [lines[i] for i in order]
    junk = ["@HD\tVN:1.6\tSO:unsorted", "@SQ\tSN:refseq\tLN:%d" % genome,
            "unmapped1\t4\t*\t0\t0\t*\t*\t0\t0\tACGT\tIIII", "unmapped2\t77\t*\t0\t0\t*\t*\t0\t0\tACGT\tIIII"]
    if head_junk:
        lines = junk[:2] + lines[:10] + junk[2:] + lines[10:]
    if tail_junk:
        lines = lines + junk
    with open(os.path.join(idir, "smp_alignment.sam"), "w") as f:
        f.write("\n".join(lines) + "\n")
    return ref


def run_sam2pb(binary, root, *, min_af="0.05", min_depth=3, min_phred=20, max_reads=1000000000, threads=4, env=None):
    args = [binary, "sam2PB", "-T", str(threads), "-i", "unused.pb", "-p", "smp", "-f", "ref.fa", "-d", "ds", "-m", str(max_reads),
            "-a", min_af, "-c", str(min_depth), "-q", str(min_phred)]
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run(args, cwd=root, capture_output=True, text=True, timeout=600, env=e)
    assert r.returncode == 0, (r.stdout[-1000:], r.stderr[-1000:])
    with open(os.path.join(root, "intermediate", "ds", "smp_reads.pb"), "rb") as f:
        return f.read()


def decode(pb):
    data = formats.SamSam()
    data.ParseFromString(pb)
    reads = [(r.read, r.start_idx, r.content, r.degree) for r in data.reads]
    rev = {c.column_name: list(c.input_columns) for c in data.reverse_columns}
    return reads, rev


@needs_ref
@pytest.mark.parametrize("seed,kw", [(1, {}), (2, dict(min_af="0.2", min_depth=8)), (3, dict(min_phred=0, min_depth=0, min_af="0.0")),
                                     (4, dict(threads=1))])
def test_sam2pb_matches_reference_binary(seed, kw, tmp_path):
    a, b = str(tmp_path / "ref"), str(tmp_path / "ours")
    make_sam_workspace(a, seed, tail_junk=False)
    make_sam_workspace(b, seed, tail_junk=False)
    want_reads, want_rev = decode(run_sam2pb(REF, a, **kw))
    got_reads, got_rev = decode(run_sam2pb(OURS, b, **kw))
    assert [r[1:] for r in got_reads] == [r[1:] for r in want_reads]          # start, content, degree, in order
    assert len(got_rev) == len(want_rev) == len(got_reads)
    for (gn, s, c, d), (wn, _, _, _) in zip(got_reads, want_reads):
        suffix = f"_READ_{s}_{s + len(c) - 1}_{d}"
        assert gn.endswith(suffix) and wn.endswith(suffix)
        g_group, w_group = sorted(got_rev[gn]), sorted(want_rev[wn])
        assert g_group == w_group and len(g_group) == d
        assert gn[: -len(suffix)] in g_group and wn[: -len(suffix)] in w_group


def test_output_is_read_by_the_loader_and_deterministic(tmp_path):
    """The product's own reader (wepp_reads_load) takes the file; two runs with different thread counts give
    identical bytes (file-order parsing, stable sort)."""
    a, b = str(tmp_path / "a"), str(tmp_path / "b")
    ref = make_sam_workspace(a, 7)
    make_sam_workspace(b, 7)
    pa = run_sam2pb(OURS, a, threads=1)
    pb_ = run_sam2pb(OURS, b, threads=5)
    assert pa == pb_
    rs = wio.parse_reads(pa, ref)
    reads, rev = decode(pa)
    assert rs.n_reads == len(reads)
    assert rs.names == [r[0] for r in reads]
    assert rs.start.tolist() == [r[1] for r in reads]
    assert rs.degree.tolist() == [r[3] for r in reads]
    assert rs.reverse_merge == rev
    assert all("_" not in r[2] for r in reads)          # gaps are turned into N (sam2pb.cpp:308-311)


def test_header_lines_anywhere(tmp_path):
    """Header / unmapped lines at the top and in the middle lose no reads (the knowing fix of the
    return-vs-continue quirk)."""
    a, b = str(tmp_path / "a"), str(tmp_path / "b")
    make_sam_workspace(a, 9, tail_junk=True, head_junk=False)
    make_sam_workspace(b, 9, tail_junk=False, head_junk=True)
    assert decode(run_sam2pb(OURS, a)) == decode(run_sam2pb(OURS, b))


def test_subsample_and_errors(tmp_path):
    a = str(tmp_path / "a")
    make_sam_workspace(a, 5, n_templates=10)
    reads, rev = decode(run_sam2pb(OURS, a, max_reads=20, env={"WEPP_SEED": "1"}))
    assert sum(r[3] for r in reads) == 20 and sum(len(v) for v in rev.values()) == 20
    again, _ = decode(run_sam2pb(OURS, a, max_reads=20, env={"WEPP_SEED": "1"}))
    assert again == reads
    os.remove(os.path.join(a, "intermediate", "ds", "smp_alignment.sam"))
    r = subprocess.run([OURS, "sam2PB", "-p", "smp", "-f", "ref.fa", "-d", "ds"], cwd=a, capture_output=True, text=True)
    assert r.returncode == 1 and "Zero reads" in r.stderr           # sam2pb.cpp:334-337
