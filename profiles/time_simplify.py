"""device time of disco_gpu_simplify on the config-2 graph (stage times with DISCO_SIMPLIFY_TRACE=1)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from disco_b200 import gpu, host, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
rs = synth.single_genome(n, 150, 30.0, seed=2)
packed, lens = host.pack_codes(rs.codes, rs.off, 8)
g = gpu.GpuBuildGraph(0)
g.load_reads(packed, lens)
g.build_graph(50, 4)
for _ in range(3):
    e, inner, st = g.simplify(50)
    print(len(e), len(inner), st, flush=True)
