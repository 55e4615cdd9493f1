#!/usr/bin/env python
"""Development aid for ncu: set up the bench shard, then N plain wepp_place(0, 0) calls.  usage: dev_one.py [scale] [n]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wepp_b200 import synth
from wepp_b200.placement import Placer
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
arena = synth.make_arena(max(int(8_000_000 * scale), 1000), 29903, synth.SEED)
reads = synth.make_reads(arena, max(int(1_250_000 * scale), 256), synth.SEED)
p = Placer(0)
p.set_arena(arena)
p.set_reads(reads)
for _ in range(n):
    p.place(0, 0)
print(p.stats())
p.close()
