"""Multi-GPU BuildGraph, one process per GPU (torch.distributed / NCCL is the plumbing).

Two partitionings, as in the reference: Mode A = ShardedBuildGraph (BuildGraphMPI: everything replicated, queries split),
described next, and Mode B = KeyShardedBuildGraph (BuildGraphMPIRMA: the hash table split by key, reached one-sidedly),
described at that class.

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


def _share_stream(g, t):
    """The library launches on the context's stream, torch / NCCL on torch's current stream: the drivers below interleave
    the two without host synchronisation, so both must be ONE stream.  Called by the constructors; a caller that switches
    torch streams afterwards has to call g.set_stream() again."""
    dev = getattr(t, "device", None)
    if dev is not None and dev.type == "cuda" and hasattr(g, "set_stream"):
        g.set_stream(torch.cuda.current_stream(dev).cuda_stream)


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


def allreduce_keys(keys: torch.Tensor, world: int, group=None, g=None):
    """MIN over ranks of the containment keys.  Only a few percent of the reads are contained, so the ranks exchange the
    (read, key) pairs that are set -- an all-gather of a few MB -- instead of all-reducing u64[n]; dense when many are.
    g: the context, whose kernels compact the keys and take the minima (without it: the same in torch ops)."""
    n = keys.numel()
    compact = getattr(g, "compact_keys", None) if keys.is_cuda else None
    cap = max(n // 8, 1024)
    if compact is not None:
        mine = torch.empty((cap, 2), dtype=torch.int64, device=keys.device)
        cnt_here = compact(mine.data_ptr(), cap)
    else:
        sel = (keys != -1).nonzero(as_tuple=True)[0]
        cnt_here = sel.numel()
    cnt = torch.tensor([cnt_here], dtype=torch.int64, device=keys.device)
    every = [torch.empty_like(cnt) for _ in range(world)]
    dist.all_gather(every, cnt, group=group)
    most = max(int(c[0]) for c in every)
    if most > cap:
        return allreduce_unsigned_min(keys, group)
    if most == 0:
        return keys
    if compact is not None:
        if cnt_here < most:
            mine[cnt_here:most, 0] = n          # padding: a read id outside the set
        allp = torch.empty((world, most, 2), dtype=torch.int64, device=keys.device)
        dist.all_gather_into_tensor(allp.view(-1), mine[:most].reshape(-1), group=group)
        g.apply_keys(allp.data_ptr(), world * most)
        return keys
    sign = torch.iinfo(torch.int64).min
    pair = torch.empty((2, most), dtype=torch.int64, device=keys.device)
    pair[0].zero_()
    pair[1].fill_(-1)                       # padding: the all-ones sentinel, loses against everything
    pair[0, :cnt_here] = sel
    pair[1, :cnt_here] = keys[sel]
    allp = torch.empty((world, 2, most), dtype=torch.int64, device=keys.device)
    dist.all_gather_into_tensor(allp.view(-1), pair.view(-1), group=group)
    keys.bitwise_xor_(sign)                 # signed order == unsigned order
    keys.scatter_reduce_(0, allp[:, 0, :].reshape(-1), allp[:, 1, :].reshape(-1).bitwise_xor(sign), reduce="amin")
    keys.bitwise_xor_(sign)
    return keys


def exchange_rowinfo(rowinfo: torch.Tensor, n: int, rank: int, world: int, group=None):
    """Every rank ends up with the row infos of all reads.  A rank's own entries are final and the others' are still zero,
    so an all-reduce(SUM) does it; with equal ranges an in-place all-gather moves half the bytes."""
    if n % world == 0:
        per = n // world
        dist.all_gather_into_tensor(rowinfo, rowinfo[rank * per:(rank + 1) * per], group=group)
    else:
        dist.all_reduce(rowinfo, op=dist.ReduceOp.SUM, group=group)


def balanced_bounds(keys: torch.Tensor, world: int, block: int = 4096):
    """Read-id bounds [world + 1] of contiguous ranges that hold (to within one block) the same number of NON-contained
    reads each.  The edge pass and the reduction only work on those, and which reads are contained is anything but uniform
    in the id: a duplicate is contained by its first copy, so low ids survive more often (80 M reads, 30x: the first eighth
    of the ids keeps 9 % more reads than the average eighth).  keys = the reduced containment keys (all ones = not
    contained), identical on every rank, hence identical bounds without an exchange.  One small device-to-host read."""
    n = keys.numel()
    nb = (n + block - 1) // block
    alive = keys == -1
    full = (n // block) * block
    counts = torch.zeros(nb, dtype=torch.int64, device=keys.device)
    if full:
        counts[:n // block] = alive[:full].view(-1, block).sum(dim=1, dtype=torch.int64)
    if full < n:
        counts[-1] = alive[full:].sum(dtype=torch.int64)
    cum = counts.cumsum(0)
    targets = (cum[-1] * torch.arange(1, world, device=keys.device, dtype=torch.int64)) // world
    idx = torch.searchsorted(cum, targets, right=False).tolist()    # first block at which the running count reaches the target
    bounds = [0]
    for i in idx:
        bounds.append(max(bounds[-1], min(n, (int(i) + 1) * block)))
    bounds.append(n)
    return bounds


def exchange_rowinfo_ranges(rowinfo: torch.Tensor, bounds, rank: int, world: int, group=None):
    """exchange_rowinfo for ranges of different sizes: every rank contributes its range padded to the longest one (an
    equal-sized all-gather through a scratch buffer) instead of all-reducing u64[n]."""
    lens = [bounds[r + 1] - bounds[r] for r in range(world)]
    m = max(max(lens), 1)
    tmp = torch.empty((world, m), dtype=rowinfo.dtype, device=rowinfo.device)
    if lens[rank]:
        tmp[rank, :lens[rank]].copy_(rowinfo[bounds[rank]:bounds[rank + 1]])
    dist.all_gather_into_tensor(tmp.view(-1), tmp[rank], group=group)
    for r in range(world):
        if r != rank and lens[r]:
            rowinfo[bounds[r]:bounds[r + 1]].copy_(tmp[r, :lens[r]])


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
    exchange_rowinfo(t.rowinfo(), t.rowinfo().numel(), rank, world, group)  # entries of rows owned by other ranks are zero here
    mark("rowinfo exchange")
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

    def __init__(self, g, rank: int, world: int, group=None, tensors=None, parts: int = None):
        self.g, self.rank, self.world, self.group = g, rank, world, group
        self.t = tensors or GpuTensors(g, torch.device("cuda", torch.cuda.current_device()))
        _share_stream(g, self.t)
        if parts is None:   # more, smaller parts as the gathered volume grows: only the last part's gather is exposed
            import os
            parts = int(os.environ.get("DISCO_PARTS", "4" if world <= 2 else "8"))
        self.parts = max(1, parts)
        self.big = None           # gathered adjacency, kept across calls
        self.comm = None

    def build_graph(self, min_overlap: int, max_edge_per_kmer: int = 4):
        g, n = self.g, self.g.n
        lo, hi = partition(n, self.rank, self.world)
        g.begin(min_overlap, max_edge_per_kmer)
        g.phase_table(False)
        g.phase_contained(lo, hi)
        allreduce_keys(self.t.keys(), self.world, self.group, self.g)
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
        exchange_rowinfo(t.rowinfo(), g.n, rank, world, self.group)  # entries of rows owned by other ranks are zero here
        for w in works:
            w.wait()
        main.wait_stream(self.comm)
        g.use_rows(self.big.data_ptr(), region)
        return int(meta[0])


class KeyShardedBuildGraph:
    """Mode B driver (BASELINE config 5; the partitioning of BuildGraphMPIRMA: src/BuildGraphMPIRMA/src/HashTable.cpp
    keeps a slice of the table per rank and fetches remote buckets with one-sided MPI_Get).

    Reads are replicated; the hash table is split by key (shard = mulhi(fingerprint, world)) and the adjacency by
    query range -- neither is ever gathered, so the per-GPU footprint of both shrinks with the number of GPUs:
      table build  : every rank scans all reads, inserts only its own keys (no communication), sets all filter bits
      probes       : the kernels read remote buckets through NVLink peer pointers (CUDA IPC mappings of the shards)
      containment  : keys u64[n] all-reduce(MIN), as in Mode A
      reduction    : row info u64[n] all-reduce(SUM); neighbours' rows are read from their owner through NVLink
    What the host exchanges: the IPC handles (64 bytes per rank, once per allocation), the two all-reduces, and the
    barriers that order the phases across ranks."""

    def __init__(self, g, rank: int, world: int, group=None, tensors=None, symmetric=None, alloc=None, shard_table: bool = True,
                 balance: bool = True):
        """balance: cut the edge pass and the reduction into ranges with equal numbers of non-contained reads
        (balanced_bounds) instead of equal numbers of reads.
        symmetric: allocate the table shards and the adjacency as torch symmetric memory (CUDA VMM allocations with
        2 MB pages, mapped into every peer at rendezvous) instead of exporting the library's cudaMalloc buffers through
        legacy CUDA IPC handles.  Measured on B200: through legacy IPC mappings, random remote reads collapse (8x
        slower) once the remote footprint exceeds 1-2 GB -- the requester's TLB reach for those mappings.  Default:
        symmetric when available (env DISCO_SYMM=0 forces IPC handles)."""
        import os
        from . import gpu as _gpu
        self.g, self.rank, self.world, self.group = g, rank, world, group
        self.t = tensors or GpuTensors(g, torch.device("cuda", torch.cuda.current_device()))
        _share_stream(g, self.t)
        self.MEM_TABLE, self.MEM_ROWS, self.HANDLE = _gpu.MEM_TABLE, _gpu.MEM_ROWS, _gpu.IPC_HANDLE_BYTES
        self.DiscoError = _gpu.DiscoError
        if symmetric is None:
            symmetric = os.environ.get("DISCO_SYMM", "1") != "0" and self.t.device.type == "cuda"
        self.symmetric = symmetric
        self.alloc = alloc                   # (n_u64) -> (int64 tensor, [peer pointers]); default: torch symmetric memory
        self._symm = {}                      # which -> (tensor, peer pointers, handle)
        # shard_table=False: the hybrid of the two modes -- table (and reads) replicated and probed locally as in Mode A,
        # adjacency partitioned by query range and read through peer pointers as in Mode B: nothing is all-gathered
        self.shard_table = shard_table
        self.balance = balance
        self.bounds = None                   # read-id bounds of the last run's edge pass / reduction
        if shard_table:
            g.set_shard(world, rank)
        else:
            g.set_partition(world, rank, False)

    def _symm_buffer(self, which, n_u64):
        """Symmetric int64 buffer of n_u64 words for `which`, (re)allocated collectively when the size changes."""
        cur = self._symm.get(which)
        if (cur is None or cur[0].numel() != n_u64) and self.alloc is not None:
            t, ptrs = self.alloc(n_u64)
            cur = (t, [int(p) for p in ptrs], None)
            self._symm[which] = cur
        if cur is None or cur[0].numel() != n_u64:
            import torch.distributed._symmetric_memory as symm
            self._symm.pop(which, None)
            t = symm.empty(n_u64, dtype=torch.int64, device=self.t.device)
            grp = self.group if self.group is not None else dist.group.WORLD
            try:
                h = symm.rendezvous(t, grp)
            except Exception:
                symm.enable_symm_mem_for_group(grp.group_name)   # older torch: groups must opt in first
                h = symm.rendezvous(t, grp)
            cur = (t, [int(p) for p in h.buffer_ptrs], h)
            self._symm[which] = cur
        return cur

    def _barrier(self):
        """Every rank's queued device work is finished when this returns (host-level: the phases are milliseconds)."""
        self.g.sync()
        flag = torch.zeros(1, dtype=torch.int64, device=self.t.device)
        dist.all_reduce(flag, group=self.group)
        flag.item()

    def _attach(self, which, bounds=None):
        """All-gather the IPC handles of `which` and map the peers' buffers."""
        mine = torch.frombuffer(bytearray(self.g.export_mem(which)), dtype=torch.uint8).to(self.t.device)
        every = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(every, mine, group=self.group)
        self.g.import_peers(which, [bytes(h.cpu().numpy().tobytes()) for h in every], bounds)

    def build_graph(self, min_overlap: int, max_edge_per_kmer: int = 4):
        g, n = self.g, self.g.n
        lo, hi = partition(n, self.rank, self.world)
        bounds = [partition(n, r, self.world)[0] for r in range(self.world)] + [n]
        g.begin(min_overlap, max_edge_per_kmer)
        if self.shard_table:
            if self.symmetric:
                words = g.table_words()
                t, ptrs, _ = self._symm_buffer(self.MEM_TABLE, words)
                if g.dev_table() != t.data_ptr():
                    g.adopt_buffer(self.MEM_TABLE, t.data_ptr(), words)
                g.import_peer_ptrs(self.MEM_TABLE, ptrs)
            g.phase_table(False)
            if not self.symmetric:
                self._attach(self.MEM_TABLE)
            self._barrier()                         # every shard complete before anybody probes it
        else:
            g.phase_table(False)                    # the whole table, on every GPU
        g.phase_contained(lo, hi)
        allreduce_keys(self.t.keys(), self.world, self.group, self.g)
        g.phase_finish_contained()
        if self.balance:                            # from here on a rank's share is counted in non-contained reads
            bounds = balanced_bounds(self.t.keys(), self.world)
            lo, hi = bounds[self.rank], bounds[self.rank + 1]
        self.bounds = bounds
        if self.shard_table:
            self._barrier()                         # everybody done probing before the shards are rebuilt
            g.phase_table(True)                     # own shard only: 1/world of the single-GPU cost
            self._barrier()
        elif self.world <= 2:
            g.phase_table(True)                     # (beyond two ranks the rebuild costs more than the bitmap checks it saves)
        if self.symmetric:
            self._edges_symmetric(lo, hi, n)
        else:
            g.phase_edges(lo, hi)               # rows of the own query range stay here
        meta = torch.tensor([int(g.stats()["max_degree"])], device=self.t.device, dtype=torch.int64)
        dist.all_reduce(meta, op=dist.ReduceOp.MAX, group=self.group)
        if self.balance:
            exchange_rowinfo_ranges(self.t.rowinfo(), bounds, self.rank, self.world, self.group)
        else:
            exchange_rowinfo(self.t.rowinfo(), n, self.rank, self.world, self.group)  # starts are offsets in the owner's buffer
        g.set_max_degree(int(meta[0]))
        if self.symmetric:
            g.import_peer_ptrs(self.MEM_ROWS, self._symm[self.MEM_ROWS][1], bounds)
        else:
            self._attach(self.MEM_ROWS, bounds)  # after the pass: an overflow retry may have reallocated the rows
        self._barrier()                         # every rank's rows complete before neighbours read them
        g.phase_reduce_mark(lo, hi)
        self._barrier()                         # emission reads the marks of remote neighbours
        g.phase_reduce_emit(lo, hi)
        self._barrier()                         # nobody may start overwriting rows while a peer still reads them

    def _edges_symmetric(self, lo, hi, n):
        """Edge pass into a symmetric adjacency buffer.  The buffer cannot grow under the library's feet (every rank
        must reallocate together), so an overflow on any rank makes all ranks retry with a larger one."""
        g = self.g
        cap = getattr(self, "_rows_cap", 0) or ((n + self.world - 1) // self.world) * 48 + (16 << 20)
        while True:
            t, _, _ = self._symm_buffer(self.MEM_ROWS, cap)
            if g.dev_rows()[0] != t.data_ptr():
                g.adopt_buffer(self.MEM_ROWS, t.data_ptr(), cap)
            ok = 1
            try:
                g.phase_edges(lo, hi)
            except self.DiscoError as e:
                if "too small" not in str(e):
                    raise
                ok = 0
            flag = torch.tensor([ok], dtype=torch.int64, device=self.t.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
            if int(flag[0]):
                break
            cap = cap * 3 // 2
        self._rows_cap = cap
