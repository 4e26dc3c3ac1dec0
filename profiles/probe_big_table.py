"""What one rank of an 8-GPU replicated run sees, on ONE GPU: the table of all 80 M reads, its own 10 M queries.  Used to
separate the table-size effects on the probe kernel (filter bits per record, contained candidates dropped through the
bitmap, footprint) from communication.  python profiles/probe_big_table.py [total reads] [query reads]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from disco_b200 import gpu

n = int(sys.argv[1]) if len(sys.argv) > 1 else 80_000_000
q = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
dev = torch.device("cuda", 0)
d_packed, d_lens = bench.make_packed_on_gpu(n, 2, dev, 8)
torch.cuda.synchronize()


def run(env, rebuild=False):
    for k in ("DISCO_FILTER_LOG2", "DISCO_TABLE_BUCKETS_X10"):
        os.environ.pop(k, None)
    os.environ.update(env)
    g = gpu.GpuBuildGraph(0)
    g.use_reads_device(d_packed.data_ptr(), d_lens.data_ptr(), n, 8, 150, 150)
    out = None
    for _ in range(2):
        g.begin(50, 4)
        g.phase_table(False)
        g.phase_contained(0, n)
        g.phase_finish_contained()
        if rebuild:
            g.phase_table(True)
        g.phase_edges(0, q)
        g.sync()
        st = g.stats()
        out = {k: round(st[k], 2) for k in ("ms_table_all", "ms_table_nc", "ms_edges_probe", "ms_edges_verify", "ms_edges_exact")}
        out["buckets_per_query"] = round(st["buckets_edges"] / max(st["queries_edges"], 1), 1)
        out["queries"] = st["queries_edges"]
    g.close()
    return out


for x10 in ("30", "20", "15", "10"):
    print("buckets per read x10 = %s, filter 2^28:" % x10, run({"DISCO_FILTER_LOG2": "28", "DISCO_TABLE_BUCKETS_X10": x10}), flush=True)
