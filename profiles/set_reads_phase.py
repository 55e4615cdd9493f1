import os, sys, time, numpy as np, torch
sys.path.insert(0, "/root/repo")
os.environ["WEPP_TIMING"] = "2"
import bench
from wepp_b200 import synth
from wepp_b200.placement import Placer
arena, reads = bench.workload(1.0, 0)
def pinned(a): return torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
rh = synth.Reads(pinned(reads.start), pinned(reads.end), pinned(reads.degree), pinned(reads.rm_off), pinned(reads.rm_pos), pinned(reads.rm_nuc))
p = Placer(0); p.set_arena(arena)
for i in range(3):
    print("--", file=sys.stderr); p.set_reads(rh)
