#!/usr/bin/env python
"""f4 timing: wepp_mat_load on a synthetic 8 M-node Parsimony::data file (CPU only, no GPU needed).
usage: python profiles/mat_load_timing.py [n_nodes] [--gz]   -> one JSON line"""
import gzip, json, os, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wepp_b200 import io, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else 8_000_000
gz = "--gz" in sys.argv
arena = synth.make_arena(n, 29903, synth.SEED)
is_leaf = np.ones(n, bool)
is_leaf[arena.parent[1:]] = False
ids = [f"S{v}|hap/{v}|2021-01-01" if is_leaf[v] else f"node_{v}" for v in range(n)]
t0 = time.time()
pb = io.serialize_mat(arena.parent, arena.mut_off, arena.mut_pos, arena.mut_ref, arena.mut_ref, arena.mut_nuc, ids)
t_ser = time.time() - t0
d = tempfile.mkdtemp(prefix="wepp_matload_")
path = os.path.join(d, "tree.pb.gz" if gz else "tree.pb")
with (gzip.open(path, "wb", compresslevel=1) if gz else open(path, "wb")) as f:
    f.write(pb)
out = {"n_nodes": n, "n_mutations": int(arena.mut_pos.shape[0]), "pb_bytes": len(pb), "file_bytes": os.path.getsize(path),
       "serialize_s": round(t_ser, 3), "threads_env": os.environ.get("WEPP_THREADS"), "cores": os.cpu_count(), "loads_s": []}
del pb
import ctypes as C
from wepp_b200 import _lib
lib = _lib.load()
for rep in range(3):   # the first load writes the sidecar (when enabled), the next ones read it
    h = C.c_void_p()
    t0 = time.time()
    _lib.check(lib.wepp_mat_load(path.encode(), 1, C.byref(h)))
    out["loads_s"].append(round(time.time() - t0, 3))
    nn, nm = C.c_int32(), C.c_int64()
    na, ic, cc = C.c_int32(), C.c_int64(), C.c_int64()
    _lib.check(lib.wepp_mat_dims(h, C.byref(nn), C.byref(nm), C.byref(na), C.byref(ic), C.byref(cc)))
    assert nn.value == n, (nn.value, n)
    lib.wepp_mat_free(h)
out["sidecar"] = [f for f in os.listdir(d) if f != os.path.basename(path)]
print(json.dumps(out))
for f in os.listdir(d):
    os.remove(os.path.join(d, f))
os.rmdir(d)
