"""Python mirror of the reference's buildG stage (src/BuildGraph/src/main.cpp:24-76) on top of the two C ABIs:
host (filter / numbering / packing / writers) and GPU (table, containment, overlap search, transitive reduction).

    bg = BuildGraph(min_overlap=50); bg.add_records(strings) | bg.add_file(path); res = bg.run(); bg.write(prefix)
"""
import os
import numpy as np
from . import gpu, host


class Result:
    pass


class BuildGraph:
    def __init__(self, min_overlap: int = 30, max_edge_per_kmer: int = 4, device: int = 0, threads: int = 0):
        self.min_overlap = min_overlap
        self.cap = max_edge_per_kmer
        self.device = device
        self.reads = host.Reads(min_overlap, threads)
        self.result = None

    def add_file(self, path):
        self.reads.add_file(path)

    def add_records(self, records):
        self.reads.add_records(records)

    def run(self) -> Result:
        r = self.reads.finalize()
        res = Result()
        res.n = r.n
        res.file_index = r.file_index.copy()
        res.lens = r.lens.copy()
        if r.n == 0:
            raise host.HostError("No reads found in the read files provided!")  # Dataset.cpp:138-139
        g = gpu.GpuBuildGraph(self.device)
        try:
            g.load_reads(np.ascontiguousarray(r.packed), np.ascontiguousarray(r.lens))
            g.build_graph(self.min_overlap, self.cap)
            res.crows = host.sort_contained(g.contained(), res.lens, self.min_overlap)
            res.edges = host.sort_edges(g.edges())
            res.stats = g.stats()
            self._g = g
        except Exception:
            g.close()
            raise
        self.result = res
        return res

    def close(self):
        if getattr(self, "_g", None):
            self._g.close()
            self._g = None

    # text forms, identical to the reference's files
    def edge_lines(self, with_flag=False):
        res = self.result
        out = []
        for e in res.edges:
            s, d, off = int(e["src"]), int(e["dst"]), int(e["offset"])
            sl, dl = int(res.lens[s]), int(res.lens[d])
            ovl = sl - off
            line = f"{int(res.file_index[s])}\t{int(res.file_index[d])}\t{int(e['orient'])},{ovl},0,0,{sl},{off},{sl - 1},{dl},0,{ovl - 1},NA"
            out.append(line + ",2" if with_flag else line)
        return out

    def crow_lines(self):
        res = self.result
        out = []
        for r in res.crows:
            c, k = int(r["contained"]), int(r["container"])
            l2, l1, st = int(res.lens[c]), int(res.lens[k]), int(r["start"])
            out.append(f"{int(res.file_index[c])}\t{int(res.file_index[k])}\t{int(r['orient'])},{l2},0,0,{l2},0,{l2},{l1},{st},{st + l2}")
        return out

    def simplified_lines(self, min_overlap: int = 0, min_reads: int = 5, min_len: int = 500):
        """parsimplify's output lines (OverlapGraphSimple.cpp:658-690) computed on the GPU from the edges still in HBM:
        `src \t dst \t orient,offset,length,0,0 \t (read,strand,offset)...`, ids as file indices."""
        res = self.result
        e, inner, st = self._g.simplify(min_overlap, min_reads, min_len)
        fi, lens = res.file_index, res.lens
        rid = (inner & np.uint64(0xFFFFFFFF)).astype(np.int64)
        off = ((inner >> np.uint64(32)) & np.uint64(0x7FFFFFFF)).astype(np.int64)
        strand = (inner >> np.uint64(63)).astype(np.int64)
        out = []
        for r in e:
            s, d = int(r["src"]), int(r["dst"])
            a, k = int(r["inner_start"]), int(r["n_inner"])
            tail = "".join(f"({int(fi[rid[i]])},{int(strand[i])},{int(off[i])})" for i in range(a, a + k))
            total = int(r["offset_total"])
            out.append(f"{int(fi[s])}\t{int(fi[d])}\t{int(r['orient'])},{total},{total + int(lens[d])},0,0\t{tail}")
        self.simplify_stats = st
        return out

    def write(self, prefix: str, shards: int = 1):
        """Writes the files runDisco.sh expects for -n <shards>: the edges as the reference's per-thread
        partial graphs (shard t owns a contiguous range of reads; mark flags 2 / 0 / 1, OverlapGraph.cpp:826-859), the
        contained rows in shard 0 (rows of one container stay together), the other containedReads files empty."""
        res = self.result
        os.makedirs(os.path.dirname(os.path.abspath(prefix)) or ".", exist_ok=True)
        host.write_pargraph_sharded(prefix, shards, res.edges, res.n, res.file_index, res.lens)
        for t in range(shards):
            c = res.crows if t == 0 else res.crows[:0]
            host.write_contained(f"{prefix}_{t}_containedReads.txt", c, res.file_index, res.lens)
