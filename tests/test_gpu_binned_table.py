"""The binned table build (k_table_bin / k_table_fill: records sorted by bucket range, the table filled slice by slice while
the slice is L2 resident) against the direct inserts and the oracle.  Small inputs get bins through DISCO_BIN_SLICE_KB; a
bin that overflows (skewed k-mers, or no head room) hands the build to the direct kernel on the device."""
import numpy as np
import pytest
from helpers import oracle_forms
from disco_b200 import gpu, host, synth
from disco_b200.buildgraph import BuildGraph

pytestmark = pytest.mark.gpu


def _run(rs, m, single_table=False):
    packed, lens = host.pack_codes(rs.codes, rs.off)
    g = gpu.GpuBuildGraph(0)
    g.load_reads(packed, lens)
    g.build_graph(m, 4)
    out = (gpu.sort_edges(g.edges()), np.sort(g.contained(), order=["contained"]), g.stats())
    g.close()
    return out


CASES = [
    ("single", lambda: synth.single_genome(30000, 150, 30.0, seed=51), 50),
    ("dup35", lambda: synth.dup_contained(20000, 150, 60.0, seed=52), 35),
    ("meta", lambda: synth.metagenome(30000, 12, 20000, 150, seed=53), 50),
    ("paired250", lambda: synth.paired_genome(6000, 250, seed=54), 30),
]


@pytest.mark.parametrize("single_table", [False, True], ids=["rebuild", "single_table"])
@pytest.mark.parametrize("name,make,m", CASES, ids=[c[0] for c in CASES])
def test_binned_equals_direct(name, make, m, single_table, monkeypatch):
    rs = make()
    if single_table:
        monkeypatch.setenv("DISCO_SINGLE_TABLE", "1")
    monkeypatch.setenv("DISCO_BINNED", "0")
    want = _run(rs, m)
    monkeypatch.delenv("DISCO_BINNED")
    for kb, headroom in (("64", None), ("8", None), ("64", "0")):     # 8 KB slices: hundreds of bins; head room 0: bins overflow
        monkeypatch.setenv("DISCO_BIN_SLICE_KB", kb)
        if headroom is not None:
            monkeypatch.setenv("DISCO_BIN_HEADROOM", headroom)
        got = _run(rs, m)
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
        for k in ("n_contained", "n_edges", "raw_directed_edges", "cap_fired", "multi_overlap_pairs", "one_sided_edges"):
            assert got[2][k] == want[2][k], k
        assert got[2]["kernel_launches"] > want[2]["kernel_launches"]        # the bin / fill kernels did run


def test_binned_against_oracle(monkeypatch):
    monkeypatch.setenv("DISCO_BIN_SLICE_KB", "32")
    rs = synth.dup_contained(12000, 150, 60.0, seed=23)
    records = rs.strings()
    o = oracle_forms(records, 35)
    bg = BuildGraph(min_overlap=35)
    bg.add_records(records)
    res = bg.run()
    try:
        assert bg.crow_lines() == o["crows"]
        assert res.stats["raw_directed_edges"] == o["res"].stats["raw_directed"]
        assert sorted(bg.edge_lines()) == o["edges"]
    finally:
        bg.close()


def test_skewed_kmers_overflow_a_bin(monkeypatch):
    """thousands of copies of one read: every record lands in two bins, which overflow -> the direct kernel builds the table"""
    monkeypatch.setenv("DISCO_BIN_SLICE_KB", "16")
    s = synth.single_genome(1, 150, 1.0, seed=41).strings()[0]
    other = synth.single_genome(3000, 150, 20.0, seed=42).strings()
    recs = [s] * 6000 + other
    bg = BuildGraph(min_overlap=50)
    bg.add_records(recs)
    res = bg.run()
    monkeypatch.setenv("DISCO_BINNED", "0")
    bg2 = BuildGraph(min_overlap=50)
    bg2.add_records(recs)
    res2 = bg2.run()
    try:
        assert len(res.crows) >= 5999 and np.array_equal(res.crows, res2.crows) and np.array_equal(res.edges, res2.edges)
    finally:
        bg.close(); bg2.close()


def test_binned_with_deferred_upload(monkeypatch):
    """the chunked upload bins each chunk as it arrives; the fill runs once all chunks are in"""
    import torch
    monkeypatch.setenv("DISCO_BIN_SLICE_KB", "512")
    rs = synth.single_genome(150_000, 150, 30.0, seed=11)
    want = _run(rs, 50)
    packed, lens = host.pack_codes(rs.codes, rs.off)
    hp = torch.from_numpy(np.ascontiguousarray(packed[:, :5]).view(np.int64)).pin_memory()
    hl = torch.from_numpy(lens.view(np.int16)).pin_memory()
    g = gpu.GpuBuildGraph(0)
    for _ in range(2):
        g.load_reads_async(hp.data_ptr(), hl.data_ptr(), rs.n, 5, 150, 150)
        g.build_graph(50, 4)
        assert np.array_equal(gpu.sort_edges(g.edges()), want[0]) and np.array_equal(np.sort(g.contained(), order=["contained"]), want[1])
    g.close()
