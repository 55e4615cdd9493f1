"""File formats either side of the path: the product's loaders
(libwepp_b200.so, csrc/host_io.cpp + pbwire.h) against the oracle restatement decoded by the real
protobuf runtime (oracle/formats.py), on files written by the real protobuf runtime."""
import gzip

import numpy as np
import pytest

from oracle import formats
from wepp_b200 import io as wio
from wepp_b200._lib import WeppError


def _random_newick(rng, n_leaves, with_len=True, with_labels=False):
    """Random rooted tree as a Newick string; returns (newick, number of nodes in preorder)."""
    counter = [0]

    def leaf():
        counter[0] += 1
        return f"S{counter[0]}|x/{counter[0]}"

    def fmt_len():
        if not with_len or rng.random() < 0.2:
            return ""
        return ":" + rng.choice(["0.5", "1e-05", "3", "0", "2.5E+1", "-1"])

    def build(k):
        if k == 1:
            return leaf() + fmt_len()
        parts, left = [], k
        n_child = int(rng.integers(2, 5))
        for c in range(n_child):
            if left <= 0:
                break
            take = left if c == n_child - 1 else int(rng.integers(1, left + 1))
            parts.append(build(take))
            left -= take
        label = f"inner{int(rng.integers(1000))}" if with_labels and rng.random() < 0.5 else ""
        return "(" + ",".join(parts) + ")" + label + fmt_len()

    return build(n_leaves) + ";"


def _random_mat(seed, n_leaves=40, meta=True, condensed=True, with_len=True, with_labels=False):
    rng = np.random.default_rng(seed)
    nw = _random_newick(rng, n_leaves, with_len, with_labels)
    parent, ids, _, _ = formats.parse_newick(nw)
    n = len(parent)
    data = formats.ParsimonyData()
    data.newick = nw
    is_leaf = np.ones(n, bool)
    for p in parent:
        if p >= 0:
            is_leaf[p] = False
    for v in range(n):
        ml = data.node_mutations.add()
        k = int(rng.integers(0, 5)) if v else int(rng.integers(0, 2))
        if is_leaf[v] and rng.random() < 0.3:
            k = 0
        positions = sorted(rng.integers(1, 200, size=k).tolist())
        for pos in positions:
            m = ml.mutation.add()
            m.position = int(pos) if rng.random() > 0.03 else -int(pos)
            m.ref_nuc = int(rng.integers(4))
            m.par_nuc = int(rng.integers(4))
            choices = [x for x in range(4)]
            m.mut_nuc.extend(sorted(rng.choice(choices, size=1 if rng.random() < 0.9 else 2, replace=False).tolist()))
            m.chromosome = "NC_045512v2"
            if rng.random() < 0.15:   # a second mutation at the same position: update or reversal
                m2 = ml.mutation.add()
                m2.position = m.position
                m2.ref_nuc = m.ref_nuc
                m2.par_nuc = int(m.mut_nuc[0])
                m2.mut_nuc.append(m.par_nuc if rng.random() < 0.5 else int(rng.integers(4)))
        if meta:
            md = data.metadata.add()
            md.clade_annotations.extend(["" if rng.random() < 0.7 else f"clade{int(rng.integers(9))}",
                                         "" if rng.random() < 0.7 else f"B.1.{int(rng.integers(9))}"])
    if condensed:
        leaves = [v for v in range(n) if is_leaf[v]]
        for v in rng.choice(leaves, size=min(8, len(leaves)), replace=False):
            cn = data.condensed_nodes.add()
            cn.node_name = ids[v]
            cn.condensed_leaves.extend([f"{ids[v]}_c{j}" for j in range(int(rng.integers(1, 5)))])
    return data.SerializeToString()


def _compare_mat(got: wio.MatTree, want: dict):
    assert got.parent.tolist() == want["parent"]
    assert got.ids == want["ids"]
    assert got.n_annotations == want["n_annotations"]
    for v in range(got.n_nodes):
        a, b = int(got.mut_off[v]), int(got.mut_off[v + 1])
        mine = [(int(got.mut_pos[k]), int(got.mut_ref[k]), int(got.mut_par[k]), int(got.mut_nuc[k])) for k in range(a, b)]
        theirs = [(m["pos"], m["ref"], m["par"], m["nuc"]) for m in want["muts"][v]]
        assert mine == theirs, v
        wc = list(want["clades"][v]) + [""] * (want["n_annotations"] - len(want["clades"][v]))
        assert got.clades[v] == wc[: want["n_annotations"]], v


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("uncondense", [False, True])
def test_mat_loader_matches_oracle(seed, uncondense):
    pb = _random_mat(seed, n_leaves=30 + 7 * seed, meta=seed % 3 != 2, condensed=True, with_len=seed % 2 == 0,
                     with_labels=seed % 3 == 0)
    _compare_mat(wio.parse_mat(pb, uncondense), formats.load_mat(pb, uncondense))


def test_mat_loader_gzip_by_file_name(tmp_path):
    pb = _random_mat(11)
    plain, gz = tmp_path / "tree.pb", tmp_path / "tree.pb.gz"
    plain.write_bytes(pb)
    with gzip.open(gz, "wb") as f:
        f.write(pb)
    want = formats.load_mat(pb, True)
    _compare_mat(wio.load_mat(str(plain)), want)
    _compare_mat(wio.load_mat(str(gz)), want)
    with pytest.raises(WeppError):
        wio.load_mat(str(tmp_path / "missing.pb"))
    (tmp_path / "bad.pb.gz").write_bytes(pb)          # named .gz but not gzip
    with pytest.raises(WeppError):
        wio.load_mat(str(tmp_path / "bad.pb.gz"))


def test_mat_sidecar_round_trip_and_invalidation(tmp_path, monkeypatch):
    """The flattened-tree sidecar (host_io.cpp): the second load of a file reads "<file>.wepp_flat" and gives the same
    tree (condensed nodes, metadata, masked mutations and all); a changed source, a truncated or foreign sidecar and
    WEPP_SIDECAR=0 all fall back to parsing; WEPP_SIDECAR_DIR moves it."""
    monkeypatch.delenv("WEPP_SIDECAR", raising=False)
    monkeypatch.delenv("WEPP_SIDECAR_DIR", raising=False)
    pb = _random_mat(21, n_leaves=60)
    src = tmp_path / "tree.pb.gz"
    with gzip.open(src, "wb") as f:
        f.write(pb)
    side = tmp_path / "tree.pb.gz.wepp_flat"
    for uncondense in (False, True):
        want = formats.load_mat(pb, uncondense)
        _compare_mat(wio.load_mat(str(src), uncondense), want)       # parses (and writes the sidecar the first time)
        assert side.exists()
        _compare_mat(wio.load_mat(str(src), uncondense), want)       # reads the sidecar
    good = side.read_bytes()
    # the sidecar really is what is read: a sidecar of ANOTHER tree under this source's key must show through
    pb2 = _random_mat(22, n_leaves=33)
    src2 = tmp_path / "other.pb"
    src2.write_bytes(pb2)
    wio.load_mat(str(src2))
    other = (tmp_path / "other.pb.wepp_flat").read_bytes()
    import struct
    hdr = struct.calcsize("<8sQQq")
    side.write_bytes(good[:hdr] + other[hdr:])
    _compare_mat(wio.load_mat(str(src)), formats.load_mat(pb2, True))
    # ... and a stale one must not: change the source
    side.write_bytes(good)
    pb3 = _random_mat(23, n_leaves=60)
    with gzip.open(src, "wb") as f:
        f.write(pb3)
    _compare_mat(wio.load_mat(str(src)), formats.load_mat(pb3, True))
    assert side.read_bytes() != good                                   # rewritten for the new source
    # truncated / garbage sidecars are ignored
    fresh = side.read_bytes()
    for bad in (fresh[: len(fresh) // 2], b"not a sidecar", fresh[:-8], fresh + b"12345678"):
        side.write_bytes(bad)
        _compare_mat(wio.load_mat(str(src)), formats.load_mat(pb3, True))
    # WEPP_SIDECAR=0: neither read nor written
    side.unlink()
    monkeypatch.setenv("WEPP_SIDECAR", "0")
    _compare_mat(wio.load_mat(str(src)), formats.load_mat(pb3, True))
    assert not side.exists()
    monkeypatch.delenv("WEPP_SIDECAR")
    # WEPP_SIDECAR_DIR
    d = tmp_path / "cache"
    d.mkdir()
    monkeypatch.setenv("WEPP_SIDECAR_DIR", str(d))
    _compare_mat(wio.load_mat(str(src)), formats.load_mat(pb3, True))
    assert (d / "tree.pb.gz.wepp_flat").exists() and not side.exists()
    # a directory that cannot be written to is not an error
    monkeypatch.setenv("WEPP_SIDECAR_DIR", str(tmp_path / "does" / "not" / "exist"))
    _compare_mat(wio.load_mat(str(src)), formats.load_mat(pb3, True))


@pytest.mark.parametrize("threads", ["1", "3", "16"])
def test_mat_loader_threads_and_unusual_newick(threads, monkeypatch):
    """The token scan and the per-node mutation fill run on WEPP_THREADS host threads (the reference's
    tbb::parallel_for, mutation_annotated_tree.cpp:556-596): any thread count gives the same tree — here on trees large
    enough to be cut into ranges; Newick tokens of the unusual shape ('(' after a name or after ')') take the
    one-thread token machine and still agree with the oracle's."""
    monkeypatch.setenv("WEPP_THREADS", threads)
    monkeypatch.setenv("WEPP_SIDECAR", "0")
    pb = _random_mat(31, n_leaves=9000, meta=True, condensed=True)
    _compare_mat(wio.parse_mat(pb, True), formats.load_mat(pb, True))
    for nw in ("((A:0.1,B:0.2):0.3,(C,D)x:0.5)root;", "(A(B,C),D);", "((A,B)(C,D),E);", "((A:1,B):2,C:3):4;"):
        data = formats.ParsimonyData()
        data.newick = nw
        try:
            parent, ids, _, _ = formats.parse_newick(nw)
        except Exception:
            parent = None                        # the token machine runs out of branch lengths: malformed for both
        for _ in range(len(parent) if parent else 8):
            data.node_mutations.add()
        if parent is None:
            with pytest.raises(WeppError):
                wio.parse_mat(data.SerializeToString(), False)
            continue
        got = wio.parse_mat(data.SerializeToString(), False)
        assert got.parent.tolist() == parent and got.ids == ids, nw


def test_mat_loader_rejects_bad_newick_and_truncated_files():
    data = formats.ParsimonyData()
    data.newick = "((A,B),C;"
    for _ in range(5):
        data.node_mutations.add()
    with pytest.raises(WeppError):
        wio.parse_mat(data.SerializeToString())
    data.newick = "((A,B),A);"      # duplicate id: "already in the tree", mutation_annotated_tree.cpp:868-871
    with pytest.raises(WeppError):
        wio.parse_mat(data.SerializeToString())
    pb = _random_mat(3)
    with pytest.raises(WeppError):
        wio.parse_mat(pb[: len(pb) // 2])


def test_mat_serialize_round_trip():
    """wepp_mat_serialize -> real protobuf decoder -> same tree; -> product loader -> same arrays."""
    want = formats.load_mat(_random_mat(5, condensed=False, meta=False), False)
    n = len(want["parent"])
    mo, mp, mr, mpar, mn = [0], [], [], [], []
    for v in range(n):
        for m in want["muts"][v]:
            if m["pos"] < 0 or m["nuc"] == 0:
                continue
            mp.append(m["pos"]); mr.append(m["ref"]); mpar.append(m["par"]); mn.append(m["nuc"])
        mo.append(len(mp))
    pb = wio.serialize_mat(want["parent"], mo, mp, mr, mpar, mn, want["ids"])
    back = formats.load_mat(pb, False)
    assert back["parent"] == want["parent"] and back["ids"] == want["ids"]
    got = wio.parse_mat(pb, False)
    assert got.mut_off.tolist() == mo and got.mut_pos.tolist() == mp
    assert got.mut_ref.tolist() == mr and got.mut_par.tolist() == mpar and got.mut_nuc.tolist() == mn


def _random_reads(seed, n=200, g=500):
    rng = np.random.default_rng(seed)
    ref = "".join(rng.choice(list("ACGT"), size=g))
    data = formats.SamSam()
    for i in range(n):
        length = int(rng.integers(0 if i == 0 else 1, 120))
        start = int(rng.integers(1, g - length + 1))
        content = list(ref[start - 1:start - 1 + length])
        for k in range(length):
            u = rng.random()
            if u < 0.05:
                content[k] = "N"
            elif u < 0.10:
                content[k] = str(rng.choice(list("ACGT")))
            elif u < 0.13:
                content[k] = "_"
        r = data.reads.add()
        r.read = f"q{i}_READ_{start}_{start + length - 1}_{i % 3 + 1}"
        r.start_idx = start
        r.content = "".join(content)
        r.degree = i % 3 + 1
        col = data.reverse_columns.add()
        col.column_name = r.read
        col.input_columns.extend([f"raw{i}_{j}" for j in range(i % 3 + 1)])
    extra = data.reverse_columns.add()        # a repeated key appends (sam2pb.cpp:539-544)
    extra.column_name = data.reads[1].read
    extra.input_columns.append("late")
    return ref, data.SerializeToString()


@pytest.mark.parametrize("seed", range(3))
def test_reads_loader_matches_oracle(seed, tmp_path):
    ref, pb = _random_reads(seed)
    want = formats.load_reads(pb, ref)
    path = tmp_path / "x_reads.pb"
    path.write_bytes(pb)
    for got in (wio.parse_reads(pb, ref, 1), wio.load_reads(str(path), ref, 4)):
        assert got.n_reads == len(want["reads"])
        for i, w in enumerate(want["reads"]):
            assert got.names[i] == w["read"]
            assert (int(got.start[i]), int(got.end[i]), int(got.degree[i])) == (w["start"], w["end"], w["degree"])
            a, b = int(got.rm_off[i]), int(got.rm_off[i + 1])
            assert list(zip(got.rm_pos[a:b].tolist(), got.rm_nuc[a:b].tolist())) == w["mutations"]
        assert got.reverse_merge == want["reverse_merge"]


def test_reads_loader_errors(tmp_path):
    ref, pb = _random_reads(1)
    with pytest.raises(WeppError):
        wio.load_reads(str(tmp_path / "nope.pb"), ref)
    with pytest.raises(WeppError):
        wio.parse_reads(pb, ref[:50])          # reads beyond the reference
    with pytest.raises(WeppError):
        wio.parse_reads(pb[:-7], ref)


# ---- pinned on the reference's own loader object code (oracle/_ref/wepp_ref, see oracle/Makefile) ----
import os
import subprocess

REF_CLI = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "wepp_ref")
needs_ref_cli = pytest.mark.skipif(not os.path.exists(REF_CLI), reason="oracle/_ref/wepp_ref not built (no /root/reference here)")


def _ref_loadmat(path, uncondense):
    out = subprocess.run([REF_CLI, "loadmat", str(path), "1" if uncondense else "0"], capture_output=True, text=True,
                         check=True).stdout
    nodes = {}
    order = []
    for line in out.splitlines():
        ident, par, muts, clades = line.split("\t")
        m = [tuple(int(x) for x in t.split(":")) for t in muts.split(",") if t]
        nodes[ident] = (par, m, clades.split("|")[:-1])
        order.append(ident)
    return nodes, order


@needs_ref_cli
@pytest.mark.parametrize("seed", range(4))
@pytest.mark.parametrize("uncondense", [False, True])
def test_reference_loader_builds_the_same_tree(seed, uncondense, tmp_path):
    """MAT::load_mutation_annotated_tree + uncondense_leaves, the reference's object code, on a gzipped
    file written by the real protobuf runtime.  The reference expands condensed nodes in hash-map order, so
    the comparison is by node id (parent id, mutations, annotations) and the fixture holds at most one
    condensed node that gets a fresh node_<k> id."""
    rng = np.random.default_rng(100 + seed)
    pb = _random_mat(100 + seed, n_leaves=25 + 5 * seed, meta=seed != 3, condensed=False, with_len=seed % 2 == 0,
                     with_labels=seed == 1)
    data = formats.ParsimonyData()
    data.ParseFromString(pb)
    parent, ids, _, _ = formats.parse_newick(data.newick)
    kids = {p for p in parent if p >= 0}
    leaves = [v for v in range(len(parent)) if v not in kids]
    with_muts = [v for v in leaves if len(data.node_mutations[v].mutation) > 0 and
                 formats.load_mat(pb, False)["muts"][v]]
    without = [v for v in leaves if not formats.load_mat(pb, False)["muts"][v]]
    chosen = []
    if with_muts:
        chosen.append((with_muts[0], 3))                 # renamed node_<k>, three new children
        chosen += [(v, 1) for v in with_muts[1:4]]       # single sample: plain rename
    chosen += [(v, int(rng.integers(1, 4))) for v in without[:3]]   # siblings appended to the parent
    for v, k in chosen:
        cn = data.condensed_nodes.add()
        cn.node_name = ids[v]
        cn.condensed_leaves.extend([f"{ids[v]}_c{j}" for j in range(k)])
    pb = data.SerializeToString()
    path = tmp_path / "tree.pb.gz"
    with gzip.open(path, "wb") as f:
        f.write(pb)
    ref_nodes, ref_order = _ref_loadmat(path, uncondense)
    got = wio.load_mat(str(path), uncondense)
    assert sorted(got.ids) == sorted(ref_nodes)
    for v, ident in enumerate(got.ids):
        par, muts, clades = ref_nodes[ident]
        assert (got.ids[got.parent[v]] if got.parent[v] >= 0 else "") == par, ident
        a, b = int(got.mut_off[v]), int(got.mut_off[v + 1])
        mine = [(int(got.mut_pos[k]), int(got.mut_ref[k]), int(got.mut_par[k]), int(got.mut_nuc[k])) for k in range(a, b)]
        assert mine == [tuple(np.int8(x) if i else x for i, x in enumerate(t)) for t in muts], ident
        assert [c for c in got.clades[v]] == (clades + [""] * got.n_annotations)[: got.n_annotations], ident
    if not uncondense:   # without condensed-node expansion the preorder itself is defined
        assert got.ids == ref_order
