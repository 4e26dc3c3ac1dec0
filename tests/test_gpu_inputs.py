"""The three ways reads reach the device must give the same graph: disco_gpu_load_reads (one blocking copy),
disco_gpu_load_reads_async (copy deferred into the table build, chunk by chunk on a copy stream) and
disco_gpu_use_reads_device (the caller's device buffers in place).  Uniform and mixed read lengths, compact and
full-pitch host rows, pinned and pageable memory, several runs on one context."""
import numpy as np
import pytest
from disco_b200 import gpu, host, synth

pytestmark = pytest.mark.gpu


def _result(g):
    return gpu.sort_edges(g.edges()), np.sort(g.contained(), order=["contained"])


def _same(a, b):
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def _workload(name):
    if name == "uniform150":
        return synth.single_genome(150_000, 150, 30.0, seed=11), 50
    return synth.dup_contained(120_000, read_len=150, coverage=40.0, min_len=100, seed=5), 35


@pytest.mark.parametrize("name", ["uniform150", "mixed100-150"])
def test_loaders_agree(name, monkeypatch):
    rs, m = _workload(name)
    import torch
    packed, lens = host.pack_codes(rs.codes, rs.off)          # full pitch (words of the longest read, even)
    n, wpr = packed.shape
    g = gpu.GpuBuildGraph(0)
    g.load_reads(packed, lens)
    g.build_graph(m, 4)
    want = _result(g)
    assert len(want[0]) > 0 and len(want[1]) > 0
    mn, mx = int(lens.min()), int(lens.max())
    words = (mx + 31) // 32
    compact = np.ascontiguousarray(packed[:, :words])         # the pitch bench.py uploads (5 words for 150 bp)
    for chunks in ("1", "3", "8", "16"):
        monkeypatch.setenv("DISCO_UPLOAD_CHUNKS", chunks)
        for rows in (packed, compact):
            for pinned in (False, True):
                for hints in ((0, 0), (mn, mx)):
                    hp = torch.from_numpy(rows.view(np.int64))
                    hl = torch.from_numpy(lens.view(np.int16))
                    if pinned:
                        hp, hl = hp.pin_memory(), hl.pin_memory()
                    g.load_reads_async(hp.data_ptr(), hl.data_ptr(), n, rows.shape[1], *hints)
                    g.build_graph(m, 4)
                    _same(_result(g), want)
    # device buffers in place (library pitch) and with another pitch (copied)
    stride = 1
    while stride < words:
        stride *= 2
    dev = torch.device("cuda", 0)
    full = np.zeros((n, stride), dtype=np.uint64)
    full[:, :words] = compact
    for rows in (full, compact):
        dp = torch.from_numpy(rows.view(np.int64)).to(dev)
        dl = torch.from_numpy(lens.view(np.int16)).to(dev)
        g.use_reads_device(dp.data_ptr(), dl.data_ptr(), n, rows.shape[1], mn, mx)
        g.build_graph(m, 4)
        _same(_result(g), want)
        g.build_graph(m, 4)                                    # the borrowed buffers serve a second run
        _same(_result(g), want)
    # back to an owned copy on the same context, then a differently sized batch
    g.load_reads(packed, lens)
    g.build_graph(m, 4)
    _same(_result(g), want)
    half = n // 2
    g.load_reads_async(torch.from_numpy(compact[:half].view(np.int64)).pin_memory().data_ptr(),
                       torch.from_numpy(lens[:half].view(np.int16)).pin_memory().data_ptr(), half, words, 0, 0)
    g.build_graph(m, 4)
    g2 = gpu.GpuBuildGraph(0)
    g2.load_reads(np.ascontiguousarray(packed[:half]), np.ascontiguousarray(lens[:half]))
    g2.build_graph(m, 4)
    _same(_result(g), _result(g2))
    g.close(); g2.close()


@pytest.mark.parametrize("n", [5, 3000, 400_000])
def test_device_edge_sort(n):
    """disco_gpu_sort_edges puts the reduced edges into (src, dst) order where they lie in HBM: get_edges then returns
    exactly what the host-side sort of the emission order gives -- also when a pinned sink mirrored the emission."""
    import torch
    rs = synth.single_genome(n, 150, 30.0 if n > 100 else 2.0, seed=3)
    packed, lens = host.pack_codes(rs.codes, rs.off)
    g = gpu.GpuBuildGraph(0)
    g.load_reads(packed, lens)
    sink = torch.empty((n * 2 + 16, 4), dtype=torch.int32).pin_memory()
    view = sink.numpy().view(gpu.EDGE_DTYPE).reshape(-1)
    g.set_edge_sink(sink.data_ptr(), sink.shape[0])
    g.build_graph(50, 4)
    raw = g.edges(out=view).copy()              # emission order, from the sink
    want = gpu.sort_edges(raw)
    g.sort_edges()
    got = g.edges()
    assert np.array_equal(got, want)
    assert np.array_equal(g.edges(out=view), want)   # the sink is refreshed from the device after a sort
    if n > 100:
        assert len(got) > n // 2 and not np.array_equal(raw, want)
    g.close()


def test_contained_rows_by_range():
    """disco_gpu_get_contained_range: the ranges of a partition hand out every contained row exactly once"""
    rs = synth.dup_contained(60_000, read_len=150, coverage=40.0, min_len=100, seed=6)
    packed, lens = host.pack_codes(rs.codes, rs.off)
    g = gpu.GpuBuildGraph(0)
    g.load_reads(packed, lens)
    g.build_graph(35, 4)
    allrows = np.sort(g.contained(), order=["contained"])
    assert len(allrows) > 1000
    buf = np.zeros(len(allrows), dtype=gpu.CROW_DTYPE)
    for bounds in ([0, rs.n], [0, 1, 4096, 4097, 30_000, rs.n], [0, 0, rs.n, rs.n]):
        parts = [g.contained_range_into(buf, a, b).copy() for a, b in zip(bounds, bounds[1:])]
        for (a, b), p in zip(zip(bounds, bounds[1:]), parts):
            assert ((p["contained"] >= a) & (p["contained"] < b)).all()
        assert np.array_equal(np.sort(np.concatenate(parts), order=["contained"]), allrows)
    with pytest.raises(gpu.DiscoError):
        g.contained_range_into(buf[:10], 0, rs.n)        # capacity too small: reported
    g.close()
