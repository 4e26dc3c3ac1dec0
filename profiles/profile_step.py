"""One hot-path pass for ncu (profiles/README.md has the commands).  Usage: python profiles/profile_step.py [reads] [warmup]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from disco_b200 import gpu, host, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rs = synth.single_genome(n, 150, 30.0, seed=2)
packed, lens = host.pack_codes(rs.codes, rs.off, 8)
g = gpu.GpuBuildGraph(0)
g.load_reads(packed, lens)
for _ in range(warm + 1):
    g.build_graph(50, 4)
print(g.stats())
