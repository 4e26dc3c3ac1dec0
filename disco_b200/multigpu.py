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


class GpuTensors:
    """Adapter: the context's device buffers as torch tensors (zero copy)."""

    def __init__(self, g, device):
        self.g, self.device = g, device

    def keys(self):
        return torch.as_tensor(_DevArray(self.g.dev_contained_keys(), self.g.n), device=self.device)

    def rowinfo(self):
        return torch.as_tensor(_DevArray(self.g.dev_rowinfo(), self.g.n), device=self.device)

    def rows(self):
        ptr, used = self.g.dev_rows()
        return torch.as_tensor(_DevArray(ptr, used), device=self.device) if used else torch.empty(0, dtype=torch.int64, device=self.device)

    def adopt_rows(self, t):
        self.g.adopt_rows(t.data_ptr(), t.numel())


def partition(n: int, rank: int, world: int):
    """Contiguous read-id block of a rank (BuildGraphMPI/src/OverlapGraph.cpp:524-529)."""
    return (rank * n) // world, ((rank + 1) * n) // world


def allreduce_unsigned_min(keys: torch.Tensor, group=None):
    """In-place MIN over ranks of u64 values stored in an int64 tensor: flipping the sign bit makes signed order equal
    unsigned order (the all-ones 'not contained' sentinel must lose against every real key)."""
    sign = torch.iinfo(torch.int64).min
    keys.bitwise_xor_(sign)
    dist.all_reduce(keys, op=dist.ReduceOp.MIN, group=group)
    keys.bitwise_xor_(sign)
    return keys


def exchange_adjacency(local_rows: torch.Tensor, max_degree: int, rebase, rowinfo: torch.Tensor, rank: int, world: int, group=None):
    """All ranks end up with the concatenation (rank order) of everybody's rows and a row-info array that points into
    it.  `rebase(base)` must add `base` to the start field of this rank's row-info entries before the SUM."""
    dev = local_rows.device
    meta = torch.tensor([local_rows.numel(), max_degree], device=dev, dtype=torch.int64)
    allm = [torch.empty_like(meta) for _ in range(world)]
    dist.all_gather(allm, meta, group=group)
    counts = [int(m[0]) for m in allm]
    maxdeg = max(int(m[1]) for m in allm)
    bases = [0]
    for c in counts[:-1]:
        bases.append(bases[-1] + c)
    total = bases[-1] + counts[-1]
    rebase(bases[rank])
    dist.all_reduce(rowinfo, op=dist.ReduceOp.SUM, group=group)  # entries of rows owned by other ranks are zero here
    big = torch.empty(max(total, 1), dtype=torch.int64, device=dev)
    if counts[rank]:
        big[bases[rank]:bases[rank] + counts[rank]].copy_(local_rows)
    for r in range(world):
        if counts[r]:
            dist.broadcast(big[bases[r]:bases[r] + counts[r]], src=r, group=group)
    return big[:total], maxdeg, bases, counts


class ShardedBuildGraph:
    def __init__(self, g, rank: int, world: int, group=None, tensors=None):
        self.g, self.rank, self.world, self.group = g, rank, world, group
        self.t = tensors or GpuTensors(g, torch.device("cuda", torch.cuda.current_device()))

    def build_graph(self, min_overlap: int, max_edge_per_kmer: int = 4):
        g, n = self.g, self.g.n
        lo, hi = partition(n, self.rank, self.world)
        g.begin(min_overlap, max_edge_per_kmer)
        g.phase_table(False)
        g.phase_contained(lo, hi)
        allreduce_unsigned_min(self.t.keys(), self.group)
        g.phase_finish_contained()
        g.phase_table(True)
        g.phase_edges(lo, hi)
        big, maxdeg, _, _ = exchange_adjacency(self.t.rows(), int(g.stats()["max_degree"]), lambda base: g.rebase_rows(lo, hi, base),
                                               self.t.rowinfo(), self.rank, self.world, self.group)
        self.t.adopt_rows(big)
        g.set_max_degree(maxdeg)
        g.phase_reduce(lo, hi)
        g.sync()
