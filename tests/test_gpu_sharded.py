"""Mode B kernels (key-sharded table, range-partitioned adjacency) on ONE GPU: `world` contexts on the same device play
the ranks, driven in lock step from this process; the "peer" pointers are the other contexts' buffers
(disco_gpu_import_peer_ptrs) and the two all-reduces are torch ops.  Results must equal the single-table run and the
oracle.  (The multi-process / NVLink version of the same path is tests/test_multigpu_gpu.py.)"""
import numpy as np
import pytest
import torch

from disco_b200 import gpu, host, multigpu, synth
from helpers import oracle_forms

pytestmark = pytest.mark.gpu


def _run_sharded(packed, lens, m, world, cap=4):
    dev = torch.device("cuda", 0)
    gs = []
    for r in range(world):
        g = gpu.GpuBuildGraph(0)
        g.load_reads(packed, lens)
        g.set_shard(world, r)
        gs.append(g)
    ts = [multigpu.GpuTensors(g, dev) for g in gs]
    n = gs[0].n
    parts = [multigpu.partition(n, r, world) for r in range(world)]
    bounds = [p[0] for p in parts] + [n]

    def sync():
        for g in gs:
            g.sync()

    def attach(which):
        ptrs = [g.dev_table() if which == gpu.MEM_TABLE else g.dev_rows()[0] for g in gs]
        for g in gs:
            g.import_peer_ptrs(which, ptrs, bounds if which == gpu.MEM_ROWS else None)

    for g in gs:
        g.begin(m, cap)
        g.phase_table(False)
    attach(gpu.MEM_TABLE)
    sync()
    for g, (lo, hi) in zip(gs, parts):
        g.phase_contained(lo, hi)
    sync()
    sign = torch.iinfo(torch.int64).min
    keys = torch.stack([t.keys() ^ sign for t in ts]).min(dim=0).values ^ sign
    for t in ts:
        t.keys().copy_(keys)
    torch.cuda.synchronize()
    for g in gs:
        g.phase_finish_contained()
    sync()
    for g in gs:
        g.phase_table(True)
    sync()
    for g, (lo, hi) in zip(gs, parts):
        g.phase_edges(lo, hi)
    sync()
    info = torch.stack([t.rowinfo() for t in ts]).sum(dim=0)
    for t in ts:
        t.rowinfo().copy_(info)
    torch.cuda.synchronize()
    maxdeg = max(int(g.stats()["max_degree"]) for g in gs)
    for g in gs:
        g.set_max_degree(maxdeg)
    attach(gpu.MEM_ROWS)
    for g, (lo, hi) in zip(gs, parts):
        g.phase_reduce_mark(lo, hi)
    sync()
    for g, (lo, hi) in zip(gs, parts):
        g.phase_reduce_emit(lo, hi)
    sync()
    edges = [g.edges() for g in gs]
    stats = [g.stats() for g in gs]
    if sum(s["one_sided_edges"] for s in stats) == 0:   # (an edge only one endpoint sees is emitted by that endpoint)
        for e, (lo, hi) in zip(edges, parts):
            assert ((e["src"] >= lo) & (e["src"] < hi)).all()
    crows = [np.sort(g.contained(), order=["contained"]) for g in gs]
    for g in gs:
        g.close()
    return gpu.sort_edges(np.concatenate(edges)), crows, stats


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("name,make,m", [
    ("dup35", lambda: synth.dup_contained(12000, 150, 60.0, seed=23), 35),
    ("single", lambda: synth.single_genome(20000, 150, 30.0, seed=21), 50),
    ("paired250", lambda: synth.paired_genome(4000, 250, seed=25), 30),
], ids=["dup35", "single", "paired250"])
def test_key_sharded_equals_single_table(name, make, m, world):
    rs = make()
    packed, lens = host.pack_codes(rs.codes, rs.off)
    g1 = gpu.GpuBuildGraph(0)
    g1.load_reads(packed, lens)
    g1.build_graph(m, 4)
    e1, c1, s1 = gpu.sort_edges(g1.edges()), np.sort(g1.contained(), order=["contained"]), g1.stats()
    g1.close()
    e, crows, stats = _run_sharded(packed, lens, m, world)
    assert len(e1) > 0 and np.array_equal(e, e1)
    for c in crows:
        assert np.array_equal(c, c1)
    assert sum(s["raw_directed_edges"] for s in stats) == s1["raw_directed_edges"]
    assert sum(s["cap_fired"] for s in stats) == s1["cap_fired"]
    assert stats[0]["table_buckets"] >= s1["table_buckets"]          # reported for the whole table


def test_key_sharded_cap_path_matches_oracle():
    """repeats make MAX_EDGE_PER_KMER fire: the exact (sequential) kernel also walks remote shards"""
    rs = synth.repeats(4000, 150, seed=31)
    records = rs.strings()
    o = oracle_forms(records, 50)
    assert o["res"].stats["cap_fired"] > 0
    packed, lens = host.pack_codes(rs.codes, rs.off)
    g1 = gpu.GpuBuildGraph(0)
    g1.load_reads(packed, lens)
    g1.build_graph(50, 4)
    e1, s1 = gpu.sort_edges(g1.edges()), g1.stats()
    g1.close()
    e, _, stats = _run_sharded(packed, lens, 50, 2)
    assert sum(s["cap_fired"] for s in stats) == s1["cap_fired"] == o["res"].stats["cap_fired"]
    assert np.array_equal(e, e1)


def test_shard_argument_checks():
    g = gpu.GpuBuildGraph(0)
    with pytest.raises(gpu.DiscoError):
        g.set_shard(9, 0)
    with pytest.raises(gpu.DiscoError):
        g.set_shard(2, 2)
    rs = synth.single_genome(2000, 150, 20.0, seed=5)
    packed, lens = host.pack_codes(rs.codes, rs.off)
    g.load_reads(packed, lens)
    g.set_shard(2, 0)
    g.begin(50, 4)
    g.phase_table(False)
    with pytest.raises(gpu.DiscoError):
        g.phase_contained(0, g.n)            # peers not imported: fails loudly instead of probing a partial table
    g.set_shard(1, 0)                        # back to the single table
    g.build_graph(50, 4)
    assert g.counts()[1] > 0
    g.close()


def _multi(packed, lens, m, devices):
    gs = []
    for d in devices:
        g = gpu.GpuBuildGraph(d)
        g.load_reads(packed, lens)
        gs.append(g)
    gpu.build_graph_multi(gs, m, 4)
    n = gs[0].n
    parts = [multigpu.partition(n, r, len(gs)) for r in range(len(gs))]
    edges = [g.edges() for g in gs]
    for e, (lo, hi) in zip(edges, parts):
        assert ((e["src"] >= lo) & (e["src"] < hi)).all()
    crows = [np.sort(g.contained(), order=["contained"]) for g in gs]
    raw = sum(g.stats()["raw_directed_edges"] for g in gs)
    for g in gs:
        g.close()
    return gpu.sort_edges(np.concatenate(edges)), crows, raw


@pytest.mark.parametrize("world", [2, 5])
def test_single_process_driver_matches_single_table(world):
    """disco_gpu_build_graph_multi (what `buildG -g a,b,...` calls): host threads + peer pointers, here with every
    context on device 0"""
    rs = synth.dup_contained(12000, 150, 60.0, seed=41)
    packed, lens = host.pack_codes(rs.codes, rs.off)
    g1 = gpu.GpuBuildGraph(0)
    g1.load_reads(packed, lens)
    g1.build_graph(35, 4)
    e1, c1, raw1 = gpu.sort_edges(g1.edges()), np.sort(g1.contained(), order=["contained"]), g1.stats()["raw_directed_edges"]
    g1.close()
    e, crows, raw = _multi(packed, lens, 35, [0] * world)
    assert len(e1) > 0 and np.array_equal(e, e1) and raw == raw1
    for c in crows:
        assert np.array_equal(c, c1)


def test_single_process_driver_two_devices():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    rs = synth.single_genome(300_000, 150, 30.0, seed=43)
    packed, lens = host.pack_codes(rs.codes, rs.off)
    g1 = gpu.GpuBuildGraph(0)
    g1.load_reads(packed, lens)
    g1.build_graph(50, 4)
    e1, c1 = gpu.sort_edges(g1.edges()), np.sort(g1.contained(), order=["contained"])
    g1.close()
    e, crows, _ = _multi(packed, lens, 50, [0, 1])
    assert np.array_equal(e, e1) and all(np.array_equal(c, c1) for c in crows)


def test_single_process_driver_rejects_mismatched_reads():
    rs = synth.single_genome(3000, 150, 20.0, seed=44)
    packed, lens = host.pack_codes(rs.codes, rs.off)
    a, b = gpu.GpuBuildGraph(0), gpu.GpuBuildGraph(0)
    a.load_reads(packed, lens)
    b.load_reads(packed[:2000], lens[:2000])
    with pytest.raises(gpu.DiscoError):
        gpu.build_graph_multi([a, b], 50, 4)
    a.close(); b.close()
