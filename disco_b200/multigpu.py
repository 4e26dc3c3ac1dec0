"""Multi-GPU BuildGraph, one process per GPU (torch.distributed / NCCL is the plumbing).

Partitioning = the reference's BuildGraphMPI ("distributed computation": src/BuildGraphMPI/src/OverlapGraph.cpp:524-529,
:293-295): packed reads and hash table replicated on every rank, query reads split into contiguous read-id ranges.
Where the MPI code gossips `int[numReads+1]` maps every few seconds (OverlapGraph.cpp:566-575, :225-234), the offline
formulation needs exactly two exchanges:
  1. containment keys  u64[n]  all-reduce(MIN)   -> every rank derives the same contained set and rows
  2. adjacency         row info u64[n] all-reduce(SUM) + rows all-gather (in place, one slot per rank)
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

    def rows_used(self):
        return self.g.dev_rows()[1]

    def reserve_rows(self, n):
        self.g.reserve_rows(n)

    def move_rows(self, dst):
        self.g.move_rows(dst)

    def rebase_rows(self, lo, hi, base):
        self.g.rebase_rows(lo, hi, base)

    def rows_buffer(self, n):
        return torch.as_tensor(_DevArray(self.g.dev_rows()[0], n), device=self.device)

    def rows_slice(self, start, n):
        return torch.as_tensor(_DevArray(self.g.dev_rows()[0] + 8 * start, n), device=self.device)

    def set_rows_used(self, n):
        self.g.set_rows_used(n)


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


def exchange_adjacency(t, max_degree: int, lo: int, hi: int, rank: int, world: int, group=None):
    """All ranks end up with every rank's rows in one common layout -- rank r's rows at [r * slot, r * slot + count_r),
    slot = the largest count -- and a row-info array that points into it.  One in-place all-gather (each rank
    contributes its own slot of the shared buffer) instead of per-rank broadcasts; no staging copy."""
    import os
    trace = os.environ.get("DISCO_TRACE_EXCHANGE") and rank == 0
    marks = []

    def mark(name):
        if trace:
            e = torch.cuda.Event(enable_timing=True); e.record(); marks.append((name, e))
    used = t.rows_used()
    dev = t.device
    mark("start")
    meta = torch.tensor([used, max_degree], device=dev, dtype=torch.int64)
    allm = [torch.empty_like(meta) for _ in range(world)]
    dist.all_gather(allm, meta, group=group)
    counts = [int(m[0]) for m in allm]
    maxdeg = max(int(m[1]) for m in allm)
    # slot = largest count plus 2% head room: the counts jitter from step to step (warp-private slices), and a slot
    # that grows by a few entries must not trigger a reallocation of the whole buffer
    slot = max(max(counts), 1)
    slot += slot // 50 + 1024
    mark("counts")
    t.reserve_rows(world * slot)
    mark("reserve")
    t.move_rows(rank * slot)                   # this rank's rows from the front of the buffer into its slot
    t.rebase_rows(lo, hi, rank * slot)
    mark("move+rebase")
    dist.all_reduce(t.rowinfo(), op=dist.ReduceOp.SUM, group=group)  # entries of rows owned by other ranks are zero here
    mark("rowinfo all-reduce")
    buf = t.rows_buffer(world * slot)
    dist.all_gather_into_tensor(buf, buf[rank * slot:(rank + 1) * slot], group=group)
    mark("rows all-gather")
    t.set_rows_used(world * slot)
    if trace:
        torch.cuda.synchronize()
        print("exchange:", ", ".join(f"{b[0]} {a[1].elapsed_time(b[1]):.2f} ms" for a, b in zip(marks, marks[1:])),
              f"| slot {slot} entries, {world * slot * 8 / 1e9:.2f} GB gathered", flush=True)
    return maxdeg, slot, counts


class ShardedBuildGraph:
    """Mode A driver.  `parts` > 1 pipelines the adjacency exchange with the edge pass: the local query range is searched
    in `parts` pieces and the all-gather of piece c runs on the NCCL stream while piece c+1 is being searched."""

    def __init__(self, g, rank: int, world: int, group=None, tensors=None, parts: int = 4):
        self.g, self.rank, self.world, self.group = g, rank, world, group
        self.t = tensors or GpuTensors(g, torch.device("cuda", torch.cuda.current_device()))
        self.parts = max(1, parts)
        self.big = None           # gathered adjacency, kept across calls
        self.comm = None

    def build_graph(self, min_overlap: int, max_edge_per_kmer: int = 4):
        g, n = self.g, self.g.n
        lo, hi = partition(n, self.rank, self.world)
        g.begin(min_overlap, max_edge_per_kmer)
        g.phase_table(False)
        g.phase_contained(lo, hi)
        allreduce_unsigned_min(self.t.keys(), self.group)
        g.phase_finish_contained()
        # Every rank holds the whole table, so a rebuild without the contained reads costs world x the single-GPU time;
        # beyond two ranks it is cheaper to keep the one table and let the edge pass skip contained candidates via the
        # bitmap (measured: +2.6 ms probe time per 10M queries against 1.4 ms x world for the rebuild).
        if self.world <= 2:
            g.phase_table(True)
        if self.parts == 1 or not hasattr(self.t, "rows_slice"):
            g.phase_edges(lo, hi)
            maxdeg, _, _ = exchange_adjacency(self.t, int(g.stats()["max_degree"]), lo, hi, self.rank, self.world, self.group)
        else:
            maxdeg = self._edges_pipelined(lo, hi)
        g.set_max_degree(maxdeg)
        g.phase_reduce(lo, hi)
        g.sync()

    def _edges_pipelined(self, lo, hi):
        g, t, rank, world = self.g, self.t, self.rank, self.world
        dev = t.device
        if self.comm is None:
            self.comm = torch.cuda.Stream(device=dev)
        main = torch.cuda.current_stream(dev)
        bounds = [lo + (hi - lo) * c // self.parts for c in range(self.parts + 1)]
        region = 0                      # entries of the gathered buffer handed out so far
        used_prev = 0                   # where this part's rows start in the local buffer
        works = []
        for c in range(self.parts):
            g.phase_edges_part(lo, hi, bounds[c], bounds[c + 1])      # returns with the part finished (reads the cursor)
            cnt = t.rows_used() - used_prev
            meta = torch.tensor([cnt], device=dev, dtype=torch.int64)
            allm = [torch.empty_like(meta) for _ in range(world)]
            dist.all_gather(allm, meta, group=self.group)
            slot = max(max(int(m[0]) for m in allm), 1)
            # the local rows of this part become the slot [region + rank*slot, +slot) of the gathered buffer; pad the
            # local buffer to a full slot so that the (equal-sized) all-gather never reads what the next part writes
            t.reserve_rows(used_prev + slot)
            t.set_rows_used(used_prev + slot)
            t.rebase_rows(bounds[c], bounds[c + 1], region + rank * slot - used_prev)
            need = region + world * slot
            if self.big is None or self.big.numel() < need:
                est = need if c == self.parts - 1 else int(need * self.parts / (c + 1) * 1.05)
                nb = torch.empty(est, dtype=torch.int64, device=dev)
                if self.big is not None and region:
                    for w in works:
                        w.wait()
                    nb[:region].copy_(self.big[:region])
                self.big = nb
            src = t.rows_slice(used_prev, slot)
            self.comm.wait_stream(main)
            with torch.cuda.stream(self.comm):
                works.append(dist.all_gather_into_tensor(self.big[region:region + world * slot], src, group=self.group, async_op=True))
            region += world * slot
            used_prev += slot
        st = g.stats()
        meta = torch.tensor([int(st["max_degree"])], device=dev, dtype=torch.int64)
        dist.all_reduce(meta, op=dist.ReduceOp.MAX, group=self.group)
        dist.all_reduce(t.rowinfo(), op=dist.ReduceOp.SUM, group=self.group)  # entries of rows owned by other ranks are zero here
        for w in works:
            w.wait()
        main.wait_stream(self.comm)
        g.use_rows(self.big.data_ptr(), region)
        return int(meta[0])
