"""Python face of oracle/_ref/libwepp_ref.so — the reference's OWN placement object code
(shim-compiled by oracle/Makefile from /root/reference/src/WEPP/*.cpp).  TEST INFRASTRUCTURE ONLY.

The reference keeps per-process statics (site_read_map, masked sites, reference sequence:
arena.hpp:158, dataset.hpp:91,180), so a process can host ONE session; use `run_case` to get a
fresh subprocess per data set.
"""
from __future__ import annotations

import ctypes as C
import os
import pickle
import subprocess
import sys
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libwepp_ref.so")
VP = C.c_void_p
_NUC = {1: "A", 2: "C", 4: "G", 8: "T"}
_lib = None
_session_open = False


def available() -> bool:
    return os.path.exists(LIB_PATH)


def fits_in_memory(n_nodes: int) -> bool:
    """The reference's arena keeps stack_muts per node (40-byte MAT::Mutation each, arena.cpp:18-46):
    budget ~8 KB per node at SARS-CoV-2 depth plus the MAT copy."""
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                avail = int(line.split()[1]) * 1024
                return avail > n_nodes * 9000 + (4 << 30)
    except Exception:
        pass
    return False


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        lib = C.CDLL(LIB_PATH)
        lib.ref_open.restype = C.c_int
        lib.ref_arena_mut_count.restype = C.c_int64
        lib.ref_arena_stack_count.restype = C.c_int64
        lib.ref_reads_mut_count.restype = C.c_int64
        lib.ref_cartesian_map.restype = C.c_double
        lib.ref_single_read_tree.restype = C.c_int
        lib.ref_filter.restype = C.c_int
        _lib = lib
    return _lib


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def _p(a):
    return None if a is None else a.ctypes.data_as(VP)


class _quiet:
    """Silence the reference's own std::cout progress lines (fd 1) during a call."""
    def __enter__(self):
        sys.stdout.flush()
        self._saved = os.dup(1)
        self._null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self._null, 1)

    def __exit__(self, *exc):
        load().ref_flush() if hasattr(load(), "ref_flush") else None
        os.dup2(self._saved, 1)
        os.close(self._saved)
        os.close(self._null)


class Session:
    """One reference `arena` built from an (uncondensed) MAT-level tree and a read set."""

    def __init__(self, tree, reads, masked=None, threads: int = 1, workdir: str | None = None, mut_par=None):
        global _session_open
        if _session_open:
            raise RuntimeError("the reference allows one data set per process (function-local statics)")
        lib = load()
        self.lib = lib
        self._tmp = None
        if workdir is None:
            self._tmp = tempfile.TemporaryDirectory(prefix="wepp_ref_")
            workdir = self._tmp.name
        ref_seq = "".join(_NUC[int(c)] for c in tree.ref_codes[1:]).encode()
        masked = _c(masked if masked is not None else [], np.int32)
        self._keep = (_c(tree.parent, np.int32), _c(tree.mut_off, np.int64), _c(tree.mut_pos, np.int32),
                      _c(tree.mut_ref, np.uint8), _c(mut_par if mut_par is not None else tree.mut_ref, np.uint8),
                      _c(tree.mut_nuc, np.uint8), _c(reads.start, np.int32), _c(reads.end, np.int32),
                      _c(reads.degree, np.int32), _c(reads.rm_off, np.int64), _c(reads.rm_pos, np.int32),
                      _c(reads.rm_nuc, np.uint8), masked)
        k = self._keep
        cwd = os.getcwd()
        try:
            n = lib.ref_open(workdir.encode(), C.c_int32(threads), ref_seq, C.c_int32(masked.shape[0]), _p(masked),
                             C.c_int32(k[0].shape[0]), *[_p(x) for x in k[:6]], C.c_int64(k[6].shape[0]),
                             *[_p(x) for x in k[6:12]])
        finally:
            os.chdir(cwd)
        if n < 0:
            raise RuntimeError("ref_open failed")
        _session_open = True
        self.n_nodes = int(n)
        self.n_reads = int(k[6].shape[0])

    def arena(self):
        """The reference's flattened arena (arena.cpp:3-56) as arrays."""
        n = self.n_nodes
        nm, ns = int(self.lib.ref_arena_mut_count()), int(self.lib.ref_arena_stack_count())
        out = {"parent": np.zeros(n, np.int32), "source": np.zeros(n, np.int32), "leaf_count": np.zeros(n, np.int32),
               "mut_off": np.zeros(n + 1, np.int64), "mut_pos": np.zeros(nm, np.int32), "mut_ref": np.zeros(nm, np.uint8),
               "mut_nuc": np.zeros(nm, np.uint8), "st_off": np.zeros(n + 1, np.int64), "st_pos": np.zeros(ns, np.int32),
               "st_nuc": np.zeros(ns, np.uint8)}
        self.lib.ref_arena_get(*[_p(out[k]) for k in ("parent", "source", "leaf_count", "mut_off", "mut_pos", "mut_ref",
                                                     "mut_nuc", "st_off", "st_pos", "st_nuc")])
        return out

    def masked_reads(self):
        nm = int(self.lib.ref_reads_mut_count())
        off = np.zeros(self.n_reads + 1, np.int64)
        pos = np.zeros(nm, np.int32)
        nuc = np.zeros(nm, np.uint8)
        self.lib.ref_reads_get(_p(off), _p(pos), _p(nuc))
        return off, pos, nuc

    def cartesian_map(self, n_sel: int = -1, mapped=None, want_node: bool = True, want_epp: bool = True):
        r = self.n_reads if n_sel < 0 else min(n_sel, self.n_reads)
        n = self.n_nodes
        mp = np.zeros(r, np.int32)
        mu = np.zeros(r, np.int32)
        sc = np.zeros(n, np.float64) if want_node else None
        ct = np.zeros((n, 50), np.int32) if want_node else None
        dd = np.zeros(n, np.float64) if want_node else None
        eo = np.zeros(r + 1, np.int64) if want_epp else None
        cap = int(min(r * 2048 + 1, 1 << 28)) if want_epp else 0
        en = np.zeros(max(cap, 1), np.int32) if want_epp else None
        m = None if mapped is None else _c(mapped, np.uint8)
        with _quiet():
            ms = self.lib.ref_cartesian_map(C.c_int64(n_sel), _p(m), _p(mp), _p(mu), _p(sc), _p(ct), _p(dd), _p(eo),
                                            _p(en), C.c_int64(cap))
        return {"max_parsimony": mp, "multiplicity": mu, "score": sc, "counts": ct, "dist_divergence": dd,
                "epp_off": eo, "epp_nodes": None if en is None else en[: int(eo[-1])], "ms": float(ms)}

    def single_read_tree(self, read_idx: int, mapped=None):
        m = None if mapped is None else _c(mapped, np.uint8)
        mv = C.c_int32(0)
        buf = np.zeros(self.n_nodes, np.int32)
        k = self.lib.ref_single_read_tree(C.c_int64(read_idx), _p(m), C.byref(mv), _p(buf), C.c_int32(self.n_nodes))
        return int(mv.value), buf[:k].copy()

    def mutation_distance(self, cand, n_reads: int | None = None):
        cand = _c(cand, np.int32)
        r = self.n_reads if n_reads is None else n_reads
        d = np.zeros((r, cand.shape[0]), np.int32)
        self.lib.ref_mutation_distance(C.c_int32(cand.shape[0]), _p(cand), C.c_int64(r), _p(d))
        return d

    def filter(self):
        buf = np.zeros(self.n_nodes, np.int32)
        with _quiet():
            k = self.lib.ref_filter(_p(buf), C.c_int32(self.n_nodes))
        return buf[:k].copy()

    def close(self):
        if self._tmp is not None:
            self._tmp.cleanup()
            self._tmp = None


def run_case(fn_module: str, fn_name: str, *args):
    """Run `fn_module.fn_name(*args)` in a fresh interpreter (one reference session per process)
    and return its pickled result."""
    root = os.path.dirname(_HERE)
    with tempfile.NamedTemporaryFile(suffix=".pkl", delete=False) as f:
        out = f.name
    code = ("import sys, pickle; sys.path.insert(0, %r); import importlib; m = importlib.import_module(%r); "
            "r = getattr(m, %r)(*pickle.loads(%r)); pickle.dump(r, open(%r, 'wb'))"
            % (root, fn_module, fn_name, pickle.dumps(args), out))
    subprocess.check_call([sys.executable, "-c", code], cwd=root, stdout=subprocess.DEVNULL)
    try:
        return pickle.load(open(out, "rb"))
    finally:
        os.unlink(out)
