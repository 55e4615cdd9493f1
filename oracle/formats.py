"""TEST INFRASTRUCTURE — CPU restatement of the reference's file loaders, decoded with the real
protobuf runtime (google.protobuf, descriptors built at run time from the reference's schemas:
parsimony.proto:4-31, sam.proto:4-18), i.e. a wire-format implementation independent of the
product's csrc/pbwire.h.

  load_mat        <- MAT::load_mutation_annotated_tree (src/mutation_annotated_tree.cpp:522-612) with
                     create_tree_from_newick_string (:415-508), Node::add_mutation (:720-746) and, when
                     asked, Tree::uncondense_leaves (:1224-1272; condensed nodes are expanded in file
                     order here — the reference's order is that of a tbb::concurrent_unordered_map)
  load_reads      <- load_reads_from_proto (src/WEPP/sam2pb.cpp:489-549)

Parity status: the tree builder is pinned on the reference's own object code where oracle/_ref holds
the shim-compiled loader (tests/test_formats.py::test_reference_loader_*); otherwise unpinned.
Only tests/ may import this module.
"""
from __future__ import annotations

from collections import deque

from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

_T = descriptor_pb2.FieldDescriptorProto


def _field(msg, name, number, ftype, repeated=False, type_name=None):
    f = msg.field.add()
    f.name, f.number, f.type = name, number, ftype
    f.label = _T.LABEL_REPEATED if repeated else _T.LABEL_OPTIONAL
    if type_name:
        f.type_name = type_name


def _build():
    pool = descriptor_pool.DescriptorPool()
    fp = descriptor_pb2.FileDescriptorProto(name="parsimony.proto", package="Parsimony", syntax="proto3")
    m = fp.message_type.add(name="mut")
    _field(m, "position", 1, _T.TYPE_INT32)
    _field(m, "ref_nuc", 2, _T.TYPE_INT32)
    _field(m, "par_nuc", 3, _T.TYPE_INT32)
    _field(m, "mut_nuc", 4, _T.TYPE_INT32, repeated=True)
    _field(m, "chromosome", 5, _T.TYPE_STRING)
    m = fp.message_type.add(name="mutation_list")
    _field(m, "mutation", 1, _T.TYPE_MESSAGE, repeated=True, type_name=".Parsimony.mut")
    m = fp.message_type.add(name="condensed_node")
    _field(m, "node_name", 1, _T.TYPE_STRING)
    _field(m, "condensed_leaves", 2, _T.TYPE_STRING, repeated=True)
    m = fp.message_type.add(name="node_metadata")
    _field(m, "clade_annotations", 1, _T.TYPE_STRING, repeated=True)
    m = fp.message_type.add(name="data")
    _field(m, "newick", 1, _T.TYPE_STRING)
    _field(m, "node_mutations", 2, _T.TYPE_MESSAGE, repeated=True, type_name=".Parsimony.mutation_list")
    _field(m, "condensed_nodes", 3, _T.TYPE_MESSAGE, repeated=True, type_name=".Parsimony.condensed_node")
    _field(m, "metadata", 4, _T.TYPE_MESSAGE, repeated=True, type_name=".Parsimony.node_metadata")
    pool.Add(fp)
    fs = descriptor_pb2.FileDescriptorProto(name="sam.proto", package="Sam", syntax="proto3")
    m = fs.message_type.add(name="read_info")
    _field(m, "read", 1, _T.TYPE_STRING)
    _field(m, "start_idx", 3, _T.TYPE_INT32)
    _field(m, "content", 6, _T.TYPE_STRING)
    _field(m, "degree", 5, _T.TYPE_INT32)
    m = fs.message_type.add(name="column_info")
    _field(m, "column_name", 1, _T.TYPE_STRING)
    _field(m, "input_columns", 2, _T.TYPE_STRING, repeated=True)
    m = fs.message_type.add(name="sam")
    _field(m, "reads", 1, _T.TYPE_MESSAGE, repeated=True, type_name=".Sam.read_info")
    _field(m, "reverse_columns", 2, _T.TYPE_MESSAGE, repeated=True, type_name=".Sam.column_info")
    pool.Add(fs)
    get = lambda n: message_factory.GetMessageClass(pool.FindMessageTypeByName(n))
    return get("Parsimony.data"), get("Sam.sam")


ParsimonyData, SamSam = _build()

NUC_ID = {"A": 1, "C": 2, "G": 4, "T": 8, "a": 1, "c": 2, "g": 4, "t": 8, "R": 5, "Y": 10, "S": 6, "W": 9, "K": 12,
          "M": 3, "B": 14, "D": 13, "H": 11}   # everything else (incl. 'V', see :68-73) -> 15


def nuc_id(c: str) -> int:
    return NUC_ID.get(c, 15)


def _stof(s: str) -> float:
    """std::stof on the filtered branch string: longest valid float prefix."""
    import re
    m = re.match(r"[+-]?(\d+\.?\d*([eE][+-]?\d+)?|\.\d+([eE][+-]?\d+)?)", s)
    if not m:
        raise ValueError("stof: " + s)
    import numpy as np
    return float(np.float32(float(m.group(0))))


def parse_newick(newick: str):
    """create_tree_from_newick_string, :415-508 -> (parent[], id[], branch_length[], n_internal_ids)."""
    leaves, num_open, num_close = [], [], []
    branch_len = [deque() for _ in range(128)]
    level = 0
    for s in newick.split(","):
        no = nc = 0
        stop = branch_start = False
        leaf, branch = "", ""
        for c in s:
            if c == ":":
                stop, branch, branch_start = True, "", True
            elif c == "(":
                no += 1
                level += 1
                while len(branch_len) <= level:
                    branch_len.append(deque())
            elif c == ")":
                stop = True
                nc += 1
                branch_len[level].append(_stof(branch) if branch else -1.0)
                level -= 1
                branch_start = False
            elif not stop:
                leaf += c
                branch_start = False
            elif branch_start:
                if c.isdigit() or c in ".eE-+":
                    branch += c
        leaves.append(leaf)
        num_open.append(no)
        num_close.append(nc)
        branch_len[level].append(_stof(branch) if branch else -1.0)
    if level != 0:
        raise ValueError("incorrect Newick format")
    parent, ids, blen = [], [], []
    stack, curr_internal = [], 0
    for leaf, no, nc in zip(leaves, num_open, num_close):
        for _ in range(no):
            curr_internal += 1
            parent.append(stack[-1] if stack else -1)
            ids.append(f"node_{curr_internal}")
            blen.append(branch_len[level].popleft())
            level += 1
            stack.append(len(parent) - 1)
        parent.append(stack[-1])
        ids.append(leaf)
        blen.append(branch_len[level].popleft())
        for _ in range(nc):
            stack.pop()
            level -= 1
    if len(set(ids)) != len(ids):
        raise ValueError("already in the tree")
    return parent, ids, blen, curr_internal


def _add_mutation(muts: list, m: dict):
    """Node::add_mutation, :720-746 (lower_bound on position)."""
    i = 0
    while i < len(muts) and muts[i]["pos"] < m["pos"]:
        i += 1
    if i < len(muts) and muts[i]["pos"] == m["pos"]:
        if muts[i]["par"] != m["nuc"]:
            muts[i]["nuc"] = m["nuc"]
        else:
            del muts[i]
    else:
        muts.insert(i, m)


def load_mat(pb_bytes: bytes, uncondense: bool = True) -> dict:
    data = ParsimonyData()
    data.ParseFromString(pb_bytes)
    parent, ids, blen, curr_internal = parse_newick(data.newick)
    n = len(parent)
    hasmeta = len(data.metadata) > 0
    clades = [[] for _ in range(n)]
    muts = [[] for _ in range(n)]
    for idx in range(n):          # creation order == depth_first_expansion order
        if hasmeta:
            clades[idx] = list(data.metadata[idx].clade_annotations)
        for mut in data.node_mutations[idx].mutation:
            if mut.position >= 0:
                m = {"pos": mut.position, "ref": 1 << mut.ref_nuc, "par": 1 << mut.par_nuc,
                     "nuc": sum(1 << x for x in mut.mut_nuc) & 0xFF}
                if m["nuc"] != m["par"]:
                    _add_mutation(muts[idx], m)
            else:
                _add_mutation(muts[idx], {"pos": mut.position, "ref": 0, "par": 0, "nuc": 0})
    n_ann = len(clades[0]) if n else 0
    if uncondense:                # Tree::uncondense_leaves, :1224-1272
        index = {s: v for v, s in enumerate(ids)}
        added = []
        for cn in data.condensed_nodes:
            if cn.node_name not in index:
                continue
            v = index[cn.node_name]
            par = parent[v] if parent[v] >= 0 else v
            s = list(cn.condensed_leaves)
            if len(s) > 1 and muts[v]:
                del index[ids[v]]
                curr_internal += 1
                ids[v] = f"node_{curr_internal}"
                index[ids[v]] = v
                added += [(v, x, -1.0) for x in s]
            elif len(s) > 1:
                del index[ids[v]]
                ids[v] = s[0]
                index[ids[v]] = v
                added += [(par, x, blen[v]) for x in s[1:]]
            elif len(s) == 1:
                del index[ids[v]]
                ids[v] = s[0]
                index[ids[v]] = v
        for p, name, bl in added:
            parent.append(p)
            ids.append(name)
            blen.append(bl)
            clades.append([""] * n_ann)
            muts.append([])
    return {"parent": parent, "ids": ids, "branch_length": blen, "muts": muts, "clades": clades, "n_annotations": n_ann}


def load_reads(pb_bytes: bytes, reference: str) -> dict:
    """load_reads_from_proto, src/WEPP/sam2pb.cpp:489-549."""
    data = SamSam()
    data.ParseFromString(pb_bytes)
    reads = []
    for curr in data.reads:
        start = curr.start_idx
        muts = []
        for i, c in enumerate(curr.content):
            if c != reference[start + i - 1] and c != "_":
                muts.append((start + i, nuc_id(c)))
        reads.append({"read": curr.read, "start": start, "end": start + len(curr.content) - 1, "degree": curr.degree,
                      "mutations": muts})
    reverse = {}
    for inv in data.reverse_columns:
        for x in inv.input_columns:
            reverse.setdefault(inv.column_name, []).append(x)
    return {"reads": reads, "reverse_merge": reverse}
