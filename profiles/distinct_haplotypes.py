#!/usr/bin/env python
"""How many DISTINCT window-restricted haplotypes does a read window see, against the entries of its Euler list?
(CPU analysis for DESIGN.md section 6c: the scores of a read depend on a node only through the node's haplotype
restricted to the window.)   usage: python profiles/distinct_haplotypes.py [n_nodes]"""
import sys, time, numpy as np
sys.path.insert(0, "/root/repo")
from wepp_b200 import synth
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 8_000_000
arena = synth.make_arena(n, 29903, synth.SEED)
parent = arena.parent; off = arena.mut_off; pos = arena.mut_pos; ref = arena.mut_ref; nuc = arena.mut_nuc
is_leaf = np.ones(n, bool); is_leaf[parent[1:]] = False
node_of_event = np.repeat(np.arange(n), np.diff(off))
for w0 in (5000, 14000, 23000):
    w1 = w0 + 176
    sel = np.flatnonzero((pos >= w0) & (pos < w1))
    ev_nodes = node_of_event[sel]
    uniq_nodes = np.unique(ev_nodes)
    has = np.zeros(n, bool); has[uniq_nodes] = True
    # nearest ancestor with in-window events
    anc = {}
    for v in uniq_nodes.tolist():
        u = int(parent[v])
        while u >= 0 and not has[u]:
            u = int(parent[u])
        anc[v] = u
    hap = {-1: {}}
    states = {}
    entries = 0
    for v in uniq_nodes.tolist():
        h = dict(hap[anc[v]])
        for k in range(int(off[v]), int(off[v + 1])):
            p = int(pos[k])
            if w0 <= p < w1:
                if ref[k] == nuc[k]:
                    h.pop(p, None)
                else:
                    h[p] = int(nuc[k])
                entries += 1 if is_leaf[v] else 2
        hap[v] = h
        key = tuple(sorted(h.items()))
        states[key] = states.get(key, 0) + 1
    # trie over sorted mutation lists: number of distinct prefixes
    prefixes = set()
    for key in states:
        for i in range(1, len(key) + 1):
            prefixes.add(key[:i])
    print(f"window [{w0},{w1}): nodes with events {len(uniq_nodes)}, list entries {entries}, distinct restricted haplotypes {len(states)+1}, "
          f"trie nodes {len(prefixes)+1}, Euler entries of the trie ~{2*len(prefixes)}")
