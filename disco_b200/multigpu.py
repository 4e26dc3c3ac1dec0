"""Multi-GPU BuildGraph, one process per GPU (torch.distributed / NCCL is the plumbing).

Partitioning = the reference's BuildGraphMPI ("distributed computation": src/BuildGraphMPI/src/OverlapGraph.cpp:524-529,
:293-295): packed reads and hash table replicated on every rank, query reads split into contiguous read-id ranges.
Where the MPI code gossips `int[numReads+1]` maps every few seconds (OverlapGraph.cpp:566-575, :225-234), the offline
formulation needs exactly two exchanges:
  1. containment keys  u64[n]  all-reduce(MIN)   -> every rank derives the same contained set and rows
  2. adjacency         row info u64[n] all-reduce(SUM) + rows all-gather (variable length, by broadcast)
after which each rank reduces and emits the edges whose lower endpoint lies in its range (its shard of parGraph).
"""
import torch
import torch.distributed as dist


class _DevArray:
    def __init__(self, ptr, n, typestr="<i8"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def _wrap(ptr, n, device):
    return torch.as_tensor(_DevArray(ptr, n), device=device)


class ShardedBuildGraph:
    def __init__(self, g, rank: int, world: int, group=None):
        self.g, self.rank, self.world, self.group = g, rank, world, group
        self.device = torch.device("cuda", torch.cuda.current_device())

    def range(self, n):
        return (self.rank * n) // self.world, ((self.rank + 1) * n) // self.world

    def build_graph(self, min_overlap: int, max_edge_per_kmer: int = 4):
        g, n = self.g, self.g.n
        lo, hi = self.range(n)
        g.begin(min_overlap, max_edge_per_kmer)
        g.phase_table(False)
        g.phase_contained(lo, hi)
        # 1. unsigned MIN over ranks: flip the sign bit so that signed order == unsigned order
        keys = _wrap(g.dev_contained_keys(), n, self.device)
        sign = torch.iinfo(torch.int64).min
        keys.bitwise_xor_(sign)
        dist.all_reduce(keys, op=dist.ReduceOp.MIN, group=self.group)
        keys.bitwise_xor_(sign)
        g.phase_finish_contained()
        g.phase_table(True)
        g.phase_edges(lo, hi)
        # 2. adjacency exchange
        ptr, used = g.dev_rows()
        st = g.stats()
        meta = torch.tensor([used, st["max_degree"]], device=self.device, dtype=torch.int64)
        allm = [torch.empty_like(meta) for _ in range(self.world)]
        dist.all_gather(allm, meta, group=self.group)
        counts = [int(m[0]) for m in allm]
        maxdeg = max(int(m[1]) for m in allm)
        bases = [0]
        for c in counts[:-1]:
            bases.append(bases[-1] + c)
        total = bases[-1] + counts[-1]
        g.rebase_rows(lo, hi, bases[self.rank])
        rowinfo = _wrap(g.dev_rowinfo(), n, self.device)
        dist.all_reduce(rowinfo, op=dist.ReduceOp.SUM, group=self.group)  # rows of other ranks are zero here
        big = torch.empty(max(total, 1), dtype=torch.int64, device=self.device)
        if used:
            big[bases[self.rank]:bases[self.rank] + used].copy_(_wrap(ptr, used, self.device))
        for r in range(self.world):
            if counts[r]:
                dist.broadcast(big[bases[r]:bases[r] + counts[r]], src=r, group=self.group)
        g.adopt_rows(big.data_ptr(), total)
        g.set_max_degree(maxdeg)
        g.phase_reduce(lo, hi)
        g.sync()
        del big
