"""world_size-2 gloo test (CPU) of the multi-GPU host logic in disco_b200/multigpu.py: the range partition, the
unsigned all-reduce(MIN) of containment keys and the variable-length adjacency exchange, driven through a fake context
that exposes CPU tensors where the real one exposes device buffers."""
import os
import sys
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N = 1000
SENT = -1  # all ones


def _keys_for(rank, lo, hi):
    rng = np.random.default_rng(100 + rank)
    k = np.full(N, SENT, dtype=np.int64)
    idx = np.arange(N)
    # a rank can mark ANY read contained (its queries cover reads everywhere); include large (bit 50) and tiny keys
    sel = rng.random(N) < 0.3
    k[sel] = (rng.integers(lo, max(hi, lo + 1), size=int(sel.sum())).astype(np.int64) << 20) | rng.integers(0, 1 << 19, size=int(sel.sum()))
    return k


def _rows_for(rank, lo, hi):
    rng = np.random.default_rng(200 + rank)
    deg = rng.integers(0, 6, size=hi - lo)
    rows = rng.integers(1, 1 << 40, size=int(deg.sum())).astype(np.int64)
    start = np.concatenate([[0], np.cumsum(deg)[:-1]])
    info = np.zeros(N, dtype=np.int64)
    info[lo:hi] = np.where(deg > 0, (start << 20) | deg, 0)
    return rows, info, int(deg.max()) if len(deg) else 0


class FakeCtx:
    def __init__(self, rank, world):
        self.n = N
        self.rank, self.world = rank, world
        self.log = []

    def begin(self, m, cap): self.log.append("begin")
    def phase_table(self, ex): self.log.append(f"table{int(ex)}")
    def phase_contained(self, lo, hi):
        self.keys = torch.from_numpy(_keys_for(self.rank, lo, hi))
        self.log.append("contained")
    def phase_finish_contained(self): self.log.append("finish")
    def phase_edges(self, lo, hi):
        rows, info, md = _rows_for(self.rank, lo, hi)
        self.rows, self.rowinfo, self.maxdeg = torch.from_numpy(rows), torch.from_numpy(info), md
        self.log.append("edges")
    def stats(self): return {"max_degree": self.maxdeg}
    def rebase_rows(self, lo, hi, base):
        seg = self.rowinfo[lo:hi]
        nz = (seg & 0xFFFFF) != 0
        seg[nz] += base << 20
    def set_max_degree(self, d): self.final_maxdeg = d
    def phase_reduce(self, lo, hi): self.log.append("reduce")
    def sync(self): pass


class FakeTensors:
    """CPU stand-in for multigpu.GpuTensors: one growable int64 buffer with the local rows at the front."""
    def __init__(self, g):
        self.g = g
        self.device = torch.device("cpu")
        self.buf = None
    def keys(self): return self.g.keys
    def rowinfo(self): return self.g.rowinfo
    def rows_used(self): return self.g.rows.numel()
    def reserve_rows(self, n):
        self.buf = torch.zeros(n, dtype=torch.int64)
        self.buf[:self.g.rows.numel()] = self.g.rows
    def move_rows(self, dst):
        u = self.g.rows.numel()
        if dst and u:
            self.buf[dst:dst + u] = self.buf[:u].clone()
    def rebase_rows(self, lo, hi, base): self.g.rebase_rows(lo, hi, base)
    def rows_buffer(self, n): return self.buf[:n]
    def set_rows_used(self, n): self.g.adopted = self.buf[:n].clone()


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    from disco_b200 import multigpu
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = FakeCtx(rank, world)
    multigpu.ShardedBuildGraph(g, rank, world, tensors=FakeTensors(g)).build_graph(50, 4)
    q.put((rank, g.keys.numpy().copy(), g.rowinfo.numpy().copy(), g.adopted.numpy().copy(), g.final_maxdeg, g.log))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_exchange_gloo(world):
    sys.path.insert(0, ROOT)
    from disco_b200 import multigpu
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000 + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # expected: unsigned min of keys, rows concatenated in rank order, row info rebased
    parts = [multigpu.partition(N, r, world) for r in range(world)]
    assert parts[0][0] == 0 and parts[-1][1] == N and all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
    keys = np.stack([_keys_for(r, *parts[r]) for r in range(world)]).view(np.uint64).min(axis=0).view(np.int64)
    rows, infos, mds = zip(*[_rows_for(r, *parts[r]) for r in range(world)])
    slot = max(max(len(x) for x in rows), 1)
    slot += slot // 50 + 1024   # head room, see multigpu.exchange_adjacency
    for rank, k, info, adopted, md, log in res:
        assert np.array_equal(k, keys)
        assert len(adopted) == world * slot
        assert md == max(mds)
        assert log == ["begin", "table0", "contained", "finish"] + (["table1"] if world <= 2 else []) + ["edges", "reduce"]
        # every row info entry must point at that read's own row inside the gathered array
        for r in range(world):
            base = r * slot
            lo, hi = parts[r]
            loc = infos[r][lo:hi]
            for i in range(lo, hi):
                d = int(loc[i - lo] & 0xFFFFF)
                assert int(info[i] & 0xFFFFF) == d
                if d:
                    s0, s1 = int(loc[i - lo] >> 20), int(info[i] >> 20)
                    assert s1 == s0 + base
                    assert np.array_equal(adopted[s1:s1 + d], rows[r][s0:s0 + d])


# ---- Mode B (key-sharded table, range-partitioned adjacency): handle exchange, bounds, all-reduces, phase order ------
class FakeShardCtx(FakeCtx):
    def set_shard(self, world, rank): self.log.append(f"shard{rank}/{world}")
    def export_mem(self, which): return bytes([which, self.rank]) + bytes(62)   # stands in for a CUDA IPC handle
    def import_peers(self, which, handles, bounds=None):
        self.log.append(f"import{which}")
        self.imported = getattr(self, "imported", {})
        self.imported[which] = ([bytes(h) for h in handles], bounds)
    def phase_reduce_mark(self, lo, hi): self.log.append("mark")
    def phase_reduce_emit(self, lo, hi): self.log.append("emit")
    def sync(self): self.log.append("sync")


def _worker_b(rank, world, port, q, balance):
    sys.path.insert(0, ROOT)
    from disco_b200 import multigpu
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = FakeShardCtx(rank, world)
    multigpu.KeyShardedBuildGraph(g, rank, world, tensors=FakeTensors(g), balance=balance).build_graph(50, 4)
    q.put((rank, g.keys.numpy().copy(), g.rowinfo.numpy().copy(), g.final_maxdeg, g.log, g.imported))
    dist.destroy_process_group()


@pytest.mark.parametrize("balance", [False, True], ids=["by_reads", "by_survivors"])
@pytest.mark.parametrize("world", [2, 3])
def test_key_sharded_driver_gloo(world, balance):
    sys.path.insert(0, ROOT)
    from disco_b200 import multigpu
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000 + world + 7 * int(balance)
    procs = [ctx.Process(target=_worker_b, args=(r, world, port, q, balance)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    parts = [multigpu.partition(N, r, world) for r in range(world)]
    keys = np.stack([_keys_for(r, *parts[r]) for r in range(world)]).view(np.uint64).min(axis=0).view(np.int64)
    if balance:   # the edge pass and the reduction run on ranges with equal numbers of non-contained reads
        b = multigpu.balanced_bounds(torch.from_numpy(keys), world, block=4096)
        assert b[0] == 0 and b[-1] == N and all(x <= y for x, y in zip(b, b[1:]))
        parts = [(b[r], b[r + 1]) for r in range(world)]
    _, infos, mds = zip(*[_rows_for(r, *parts[r]) for r in range(world)])
    info_sum = np.sum(np.stack(infos), axis=0)     # disjoint ranges: the sum is the union; starts stay owner-local
    for rank, k, info, md, log, imported in res:
        assert np.array_equal(k, keys)
        assert np.array_equal(info, info_sum)
        assert md == max(mds)
        # a barrier (sync + all-reduce) separates: table | probing | rebuild | edge pass ... | mark | emit
        assert [x for x in log if x != "sync"] == [f"shard{rank}/{world}", "begin", "table0", "import0", "contained", "finish",
                                                   "table1", "edges", "import1", "mark", "emit"]
        order = "".join("s" if x == "sync" else "." for x in log)
        assert order.count("s") == 6
        i_mark, i_emit = log.index("mark"), log.index("emit")
        assert "sync" in log[i_mark:i_emit] and log[i_mark - 1] == "sync" and log[-1] == "sync"
        assert log[log.index("contained") - 1] == "sync" and log[log.index("table1") - 1] == "sync" and log[log.index("edges") - 1] == "sync"
        for which in (0, 1):
            handles, bounds = imported[which]
            assert handles == [bytes([which, r]) + bytes(62) for r in range(world)]   # rank order, every rank's handle
        assert list(imported[1][1]) == [p[0] for p in parts] + [N] and imported[0][1] is None


# ---- Mode B with caller-provided (symmetric) buffers: adopt + collective retry when one rank's adjacency overflows ----
class FakeSymmCtx(FakeShardCtx):
    """phase_edges fails with the library's 'too small' error until the adopted buffer reaches `need` entries"""
    class Err(RuntimeError):
        pass

    def __init__(self, rank, world, need):
        super().__init__(rank, world)
        self.need, self.table_ptr, self.rows_ptr, self.rows_cap, self.attempts = need, 0, 0, 0, 0
    def table_words(self): return 4096
    def dev_table(self): return self.table_ptr
    def dev_rows(self): return (self.rows_ptr, 0)
    def adopt_buffer(self, which, ptr, n):
        self.log.append(f"adopt{which}")
        if which == 0: self.table_ptr = ptr
        else: self.rows_ptr, self.rows_cap = ptr, n
    def import_peer_ptrs(self, which, ptrs, bounds=None):
        self.log.append(f"import{which}")
        self.imported = getattr(self, "imported", {})
        self.imported[which] = (list(ptrs), bounds)
    def phase_edges(self, lo, hi):
        self.attempts += 1
        if self.rows_cap < self.need:
            raise FakeSymmCtx.Err("adopted adjacency buffer too small: %d entries needed, %d given" % (self.need, self.rows_cap))
        super().phase_edges(lo, hi)


def _worker_symm(rank, world, port, q):
    sys.path.insert(0, ROOT)
    from disco_b200 import multigpu
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    base_cap = ((N + world - 1) // world) * 48 + (16 << 20)
    need = base_cap + 1 if rank == world - 1 else 1000        # only the last rank overflows the first buffer
    g = FakeSymmCtx(rank, world, need)
    allocs = []

    def alloc(n):      # stands in for symmetric memory: a tensor and "peer pointers" every rank agrees on
        allocs.append(n)
        t = torch.zeros(1, dtype=torch.int64).expand(n)       # numel() == n without the memory
        return t, [1000 * len(allocs) + r for r in range(world)]
    drv = multigpu.KeyShardedBuildGraph(g, rank, world, tensors=FakeTensors(g), symmetric=True, alloc=alloc)
    drv.DiscoError = FakeSymmCtx.Err
    drv.build_graph(50, 4)
    q.put((rank, g.attempts, allocs, g.log, g.imported, drv._rows_cap, base_cap))
    dist.destroy_process_group()


def test_key_sharded_symmetric_retry_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker_symm, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, attempts, allocs, log, imported, cap, base_cap in res:
        assert attempts == 2                                   # every rank repeats the pass, not only the one that overflowed
        assert allocs == [4096, base_cap, base_cap * 3 // 2] and cap == base_cap * 3 // 2
        assert log.count("adopt0") == 1 and log.count("adopt1") == 2
        assert imported[0][0] == [1000 + r for r in range(world)]          # table: first allocation
        assert imported[1][0] == [3000 + r for r in range(world)]          # adjacency: the buffer that was large enough
        assert [x for x in log if x in ("mark", "emit")] == ["mark", "emit"]


def _sparse_keys(rank, n, frac):
    rng = np.random.default_rng(300 + rank)
    k = np.full(n, SENT, dtype=np.int64)
    sel = rng.random(n) < frac
    # containers anywhere in the id space, incl. keys with the top bit set (unsigned order != signed order)
    k[sel] = (rng.integers(0, 1 << 43, size=int(sel.sum())).astype(np.int64) << 20) | rng.integers(0, 1 << 19, size=int(sel.sum()))
    return k


def _keys_worker(rank, world, port, q, n, frac):
    sys.path.insert(0, ROOT)
    from disco_b200 import multigpu
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    k = torch.from_numpy(_sparse_keys(rank, n, frac))
    multigpu.allreduce_keys(k, world)
    q.put((rank, k.numpy().copy()))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,frac", [(2, 0.03), (3, 0.05), (2, 0.0), (2, 0.5)], ids=["sparse2", "sparse3", "none", "dense"])
def test_allreduce_keys_gloo(world, frac):
    """the sparse (pairs all-gathered) and the dense (all-reduce) exchange of containment keys give the unsigned minimum"""
    n = 5000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000 + world + int(frac * 100)
    procs = [ctx.Process(target=_keys_worker, args=(r, world, port, q, n, frac)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.stack([_sparse_keys(r, n, frac) for r in range(world)]).view(np.uint64).min(axis=0).view(np.int64)
    for _, k in res:
        assert np.array_equal(k, want)


def test_balanced_bounds():
    """ranges with equal numbers of non-contained reads (to within a block), whatever the distribution of the contained ones"""
    sys.path.insert(0, ROOT)
    from disco_b200 import multigpu
    rng = np.random.default_rng(9)
    n = 1_000_003
    p = np.linspace(0.0, 0.6, n)                      # reads late in the file are contained far more often
    keys = np.where(rng.random(n) < p, rng.integers(0, 1 << 40, size=n), -1).astype(np.int64)
    alive = keys == -1
    for world in (1, 2, 3, 8):
        for block in (64, 4096):
            b = multigpu.balanced_bounds(torch.from_numpy(keys), world, block)
            assert len(b) == world + 1 and b[0] == 0 and b[-1] == n and all(x <= y for x, y in zip(b, b[1:]))
            assert all(x % block == 0 for x in b[1:-1])
            per = [int(alive[b[r]:b[r + 1]].sum()) for r in range(world)]
            assert max(per) - min(per) <= 2 * block, (world, block, per)
    # degenerate inputs: nothing contained, everything contained, fewer reads than one block
    for keys in (np.full(5000, -1, dtype=np.int64), np.zeros(5000, dtype=np.int64), np.full(10, -1, dtype=np.int64)):
        b = multigpu.balanced_bounds(torch.from_numpy(keys), 4, 4096)
        assert b[0] == 0 and b[-1] == len(keys) and all(x <= y for x, y in zip(b, b[1:]))


def _ranges_worker(rank, world, port, q, bounds, n):
    sys.path.insert(0, ROOT)
    from disco_b200 import multigpu
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    info = torch.zeros(n, dtype=torch.int64)
    info[bounds[rank]:bounds[rank + 1]] = torch.arange(bounds[rank], bounds[rank + 1]) * 1000 + rank + 1
    multigpu.exchange_rowinfo_ranges(info, bounds, rank, world)
    q.put((rank, info.numpy().copy()))
    dist.destroy_process_group()


@pytest.mark.parametrize("bounds", [[0, 300, 1000], [0, 0, 700, 1000], [0, 512, 512, 1000]], ids=["uneven2", "empty_first", "empty_middle"])
def test_exchange_rowinfo_ranges_gloo(bounds):
    n, world = bounds[-1], len(bounds) - 1
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 35500 + os.getpid() % 2000 + sum(bounds) % 97
    procs = [ctx.Process(target=_ranges_worker, args=(r, world, port, q, bounds, n)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.zeros(n, dtype=np.int64)
    for r in range(world):
        want[bounds[r]:bounds[r + 1]] = np.arange(bounds[r], bounds[r + 1]) * 1000 + r + 1
    for _, info in res:
        assert np.array_equal(info, want)
