"""GPU edge cases against the CPU oracle: tiny inputs, reads barely longer than the minimum overlap, long reads (generic
matcher, one warp per block), very high coverage (exact path), reverse-palindromic k-mers, cap values."""
import numpy as np
import pytest
from helpers import oracle_forms
from disco_b200 import gpu, host, synth
from disco_b200.buildgraph import BuildGraph

pytestmark = pytest.mark.gpu


def _both(records, m, cap=4):
    o = oracle_forms(records, m)
    bg = BuildGraph(min_overlap=m, max_edge_per_kmer=cap)
    bg.add_records(records)
    res = bg.run()
    out = (bg.crow_lines(), sorted(bg.edge_lines()), res.stats, o)
    bg.close()
    return out


def test_single_read_and_pairs():
    a = synth.single_genome(1, 120, 1.0, seed=3).strings()[0]
    crows, edges, st, o = _both([a], 40)
    assert crows == [] and edges == [] and st["n_reads"] == 1
    b = a[60:] + synth.single_genome(1, 60, 1.0, seed=4).strings()[0]      # 60-base dovetail
    brc = b[::-1].translate(str.maketrans("ACGT", "TGCA"))
    for recs in ([a, b], [b, a], [a, brc], [brc, a]):
        crows, edges, st, o = _both(recs, 40)
        assert edges == o["edges"] and len(edges) == 1
        assert crows == o["crows"] == []


def test_reads_barely_longer_than_min_overlap():
    rs = synth.single_genome(4000, 52, 40.0, seed=5)     # length 52, minOverlap 50: one search position per read
    crows, edges, st, o = _both(rs.strings(), 50)
    assert crows == o["crows"] and edges == o["edges"]
    assert st["cap_fired"] == 0


def test_long_reads_generic_matcher():
    rng = np.random.default_rng(6)
    g = synth.random_genome(rng, 60000)
    recs = []
    for _ in range(400):
        L = int(rng.integers(900, 2500))
        s = int(rng.integers(0, len(g) - L))
        r = g[s:s + L]
        if rng.random() < 0.5:
            r = synth.revcomp_codes(r)
        recs.append("".join("ACGT"[c] for c in r))
    crows, edges, st, o = _both(recs, 63)
    assert crows == o["crows"] and edges == o["edges"]


def test_very_high_coverage():
    rs = synth.single_genome(6000, 150, 400.0, seed=7)   # ~400x: one read per start and strand survives, ~500 candidates each
    recs = rs.strings()
    o = oracle_forms(recs, 50)
    bg = BuildGraph(min_overlap=50)
    bg.add_records(recs)
    res = bg.run()
    try:
        assert bg.crow_lines() == o["crows"]
        assert res.stats["max_degree"] > 128   # beyond the shared-memory queue: parked in pieces, chunked verify
        assert res.stats["cap_fired"] == o["res"].stats["cap_fired"]
        assert res.stats["raw_directed_edges"] == o["res"].stats["raw_directed"]
        rows = sorted((int(e["src"]) + 1, int(e["offset"]), int(e["dst"]) + 1, int(e["orient"])) for r in range(res.n) for e in bg._g.row(r))
        assert rows == sorted((int(e["src"]), int(e["offset"]), int(e["dst"]), int(e["orient"])) for e in o["res"].raw)
    finally:
        bg.close()


@pytest.mark.parametrize("cov", [170.0, 300.0])
def test_high_coverage_big_rows(cov):
    """130-250 candidates per read: more than the shared-memory queue holds, so the probe kernel parks them in pieces
    and the verify kernel takes its chunked path; positions rarely exceed the cap."""
    rs = synth.single_genome(8000, 150, cov, seed=17)
    recs = rs.strings()
    o = oracle_forms(recs, 50)
    bg = BuildGraph(min_overlap=50)
    bg.add_records(recs)
    res = bg.run()
    try:
        assert bg.crow_lines() == o["crows"]
        assert res.stats["max_degree"] > 128
        assert res.stats["slow_path_reads"] < res.n - len(res.crows)      # not everything fell to the exact path
        assert res.stats["cap_fired"] == o["res"].stats["cap_fired"]
        assert res.stats["raw_directed_edges"] == o["res"].stats["raw_directed"]
        rows = sorted((int(e["src"]) + 1, int(e["offset"]), int(e["dst"]) + 1, int(e["orient"])) for r in range(res.n) for e in bg._g.row(r))
        assert rows == sorted((int(e["src"]), int(e["offset"]), int(e["dst"]), int(e["orient"])) for e in o["res"].raw)
        if res.stats["cap_fired"] == 0 and res.stats["one_sided_edges"] == 0:
            assert sorted(bg.edge_lines()) == o["edges"]
    finally:
        bg.close()


def test_reverse_palindromic_kmers_even_k():
    """K = 34 (minOverlap 35): planted reverse-palindromic 34-mers at read ends.  The contained set and
    the per-read search rows must still equal the oracle's (which types them like the reference: forward only)."""
    rng = np.random.default_rng(21)
    g = synth.random_genome(rng, 12000)
    for pos in range(100, len(g) - 200, 250):
        half = g[pos:pos + 17].copy()
        g[pos + 17:pos + 34] = synth.revcomp_codes(half)
    recs = []
    for _ in range(3000):
        L = int(rng.integers(100, 151))
        if rng.random() < 0.3:
            p = int(rng.integers(0, (len(g) - 400) // 250)) * 250 + 100
            s = p if rng.random() < 0.5 else p + 34 - L
            s = max(0, min(s, len(g) - L))
        else:
            s = int(rng.integers(0, len(g) - L))
        r = g[s:s + L]
        if rng.random() < 0.5:
            r = synth.revcomp_codes(r)
        recs.append("".join("ACGT"[c] for c in r))
    o = oracle_forms(recs, 35)
    bg = BuildGraph(min_overlap=35)
    bg.add_records(recs)
    res = bg.run()
    try:
        assert bg.crow_lines() == o["crows"]
        rows = sorted((int(e["src"]) + 1, int(e["offset"]), int(e["dst"]) + 1, int(e["orient"])) for r in range(res.n) for e in bg._g.row(r))
        assert rows == sorted((int(e["src"]), int(e["offset"]), int(e["dst"]), int(e["orient"])) for e in o["res"].raw)
        assert res.stats["cap_fired"] == o["res"].stats["cap_fired"]
        # overlaps only one endpoint can see (the reference's `if / else if` typing of a palindromic record): one-sided rows,
        # and still the oracle's reduced graph edge for edge
        assert res.stats["one_sided_edges"] == o["res"].stats["one_sided_edges"]
        assert res.stats["multi_overlap_pairs"] == o["res"].stats["multi_overlap_pairs"]
        assert sorted(bg.edge_lines()) == o["edges"]
    finally:
        bg.close()


@pytest.mark.parametrize("cap", [1, 2, 8])
def test_other_cap_values(cap):
    """MAX_EDGE_PER_KMER is a compile-time constant in the reference (Common.h:62); the library takes it as an argument."""
    import ctypes as C
    from oracle import oracle
    rs = synth.repeats(2500, 150, seed=13)
    recs = rs.strings()
    bg = BuildGraph(min_overlap=50, max_edge_per_kmer=cap)
    bg.add_records(recs)
    res = bg.run()
    try:
        # per-read rows obey the cap: at most `cap` entries per k-mer position of the source read
        for r in range(0, res.n, 7):
            row = bg._g.row(r)
            L = int(res.lens[r])
            pos = np.where((row["orient"] == 3) | (row["orient"] == 2), row["offset"], L - 49 - row["offset"])
            if len(pos):
                assert np.bincount(pos.astype(np.int64)).max() <= cap
    finally:
        bg.close()


def test_buffer_overflow_retry_gives_the_same_graph(monkeypatch):
    """Adjacency and candidate buffers start from a guess (48 entries per read); a pass that overflows either is repeated
    with the size its cursors asked for.  Forced here with a guess of one entry per read, in one pass and in parts."""
    from disco_b200 import gpu, host, synth
    rs = synth.single_genome(40_000, 150, 40.0, seed=9)
    packed, lens = host.pack_codes(rs.codes, rs.off)

    def run(parts):
        g = gpu.GpuBuildGraph(0)
        g.load_reads(packed, lens)
        n = rs.n
        g.begin(50, 4)
        g.phase_table(False)
        g.phase_contained(0, n)
        g.phase_finish_contained()
        g.phase_table(True)
        b = [n * i // parts for i in range(parts + 1)]
        for i in range(parts):
            g.phase_edges_part(0, n, b[i], b[i + 1])
        g.phase_reduce(0, n)
        out = (gpu.sort_edges(g.edges()), g.stats())
        g.close()
        return out
    e0, s0 = run(1)
    monkeypatch.setenv("DISCO_ENTRIES_PER_READ", "1")
    e1, s1 = run(1)
    e3, s3 = run(3)
    assert s0["raw_directed_edges"] == s1["raw_directed_edges"] == s3["raw_directed_edges"] > 40_000 * 30
    assert np.array_equal(e0, e1) and np.array_equal(e0, e3)


def test_edge_sink_is_filled_by_the_emission_kernel():
    """disco_gpu_set_edge_sink: the kept edges land in the caller's pinned buffer while the kernel runs; get_edges into that
    buffer copies nothing, and the content equals the device-side list."""
    import torch
    rs = synth.single_genome(30_000, 150, 30.0, seed=12)
    packed, lens = host.pack_codes(rs.codes, rs.off)
    g = gpu.GpuBuildGraph(0)
    g.load_reads(packed, lens)
    g.build_graph(50, 4)
    want = gpu.sort_edges(g.edges())
    sink = torch.zeros((len(want) + 100, 4), dtype=torch.int32).pin_memory()
    g.set_edge_sink(sink.data_ptr(), sink.shape[0])
    g.load_reads(packed, lens)
    g.build_graph(50, 4)
    view = sink.numpy().view(gpu.EDGE_DTYPE).reshape(-1)
    got = g.edges(out=view)            # no copy: the kernel already wrote the sink
    assert np.array_equal(gpu.sort_edges(got.copy()), want)
    g.set_edge_sink(0, 0)
    small = torch.zeros((10, 4), dtype=torch.int32).pin_memory()
    g.set_edge_sink(small.data_ptr(), 10)                     # too small: not mirrored, the normal copy path still works
    g.load_reads(packed, lens)
    g.build_graph(50, 4)
    assert np.array_equal(gpu.sort_edges(g.edges()), want)
    g.close()
