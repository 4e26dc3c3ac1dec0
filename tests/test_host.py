"""CPU-only tests of the host side (libdisco_host.so) and of the C ABI surface of libdisco_gpu.so."""
import ctypes
import gzip
import os
import re
import subprocess
import numpy as np
import pytest
from helpers import GOLDEN, load_golden, oracle_filter, HERE
from disco_b200 import gpu, host, pack, synth

ROOT = os.path.dirname(HERE)


def test_device_primitives_on_host(tmp_path):
    """dna.cuh (hash, reverse complement, window compare, dovetail / containment geometry) against string code."""
    exe = tmp_path / "csrc_host_test"
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", str(exe), os.path.join(HERE, "csrc_host_test.cpp")], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.startswith("OK"), out.stdout + out.stderr


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    return sorted(set(re.findall(r"\b(disco_[a-z_]+)\s*\(", txt)))


def test_gpu_library_exports_every_declared_symbol():
    L = ctypes.CDLL(gpu.LIB_PATH)
    names = _declared("disco_gpu.h")
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), n
    assert sorted(gpu.EXPORTS) == names


def test_host_library_exports_every_declared_symbol():
    L = ctypes.CDLL(host.LIB_PATH)
    names = _declared("disco_host.h")
    for n in names:
        assert hasattr(L, n), n
    assert sorted(host.EXPORTS) == names


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(gpu.DiscoError):
        gpu.GpuBuildGraph(0)


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: p.split("/")[-1][:-4])
def test_filter_and_numbering_match_oracle(path):
    g = load_golden(path)
    reads, fi = oracle_filter(g["records"], g["min_overlap"])
    r = host.Reads(g["min_overlap"])
    r.add_records(g["records"])
    r.finalize()
    assert r.records == len(g["records"])
    assert r.n == len(reads)
    assert list(r.file_index) == fi
    assert list(r.lens) == [len(s) for s in reads]
    want, _ = pack.pack_codes(synth.from_strings(reads).codes, synth.from_strings(reads).off, r.packed.shape[1])
    assert np.array_equal(r.packed, want)


def test_filter_edge_cases():
    from oracle import oracle
    rng = np.random.default_rng(0)
    cases = ["ACGT" * 10, "A" * 21 + "CGT" * 3, "ACGTN" + "ACGT" * 10, "AC" * 20, "AAT" * 14, "GGGGCC" * 6,
             "ACACACACACACACACACACACACACACA" + "GATTACA" * 5, "GATTACA" * 5 + "GGGAGGGAGGGAGGGAGGGAGGGAGGGAG", "ACGT" * 7 + "A"]
    for _ in range(300):
        L = int(rng.integers(25, 80))
        p = rng.dirichlet([0.3, 0.3, 0.3, 0.3])
        cases.append("".join(rng.choice(list("ACGT"), size=L, p=p)))
    for s in cases:
        assert host.test_read(s) == oracle.test_read(s), s


def test_filter_fuzz_against_oracle():
    """The host filter decides on the 2-bit packed read with popcounts; the oracle restates Dataset::testRead on strings.
    Adversarial inputs: pattern repeats around the 50 % threshold (also straddling 32-base word boundaries), self-
    overlapping x-y-x runs, ends equal or nearly equal to the 38 filter strings, one base around 70 %, lower case, N."""
    import re
    from oracle import oracle
    rng = np.random.default_rng(123)
    pats = ["AC", "AG", "AT", "CG", "CT", "GT", "AAT", "ATA", "TAA", "AAC", "ACA", "CAA", "AAG", "AGA", "GAA", "GGGGCC",
            "A", "C", "G", "T", "ATATA", "ACAACA", "AGAGA", "TTC", "AATT"]
    filt = re.findall(r'"([ACGT]{29})"', open(os.path.join(ROOT, "disco_b200", "host", "host.cpp")).read())
    assert len(filt) == 38

    def rand_seq(L, p=None):
        return "".join(rng.choice(list("ACGT"), size=L, p=p))

    def mutate(s, rate):
        return "".join("ACGT"[rng.integers(4)] if rng.random() < rate else c for c in s)

    cases = []
    for _ in range(8000):
        L = int(rng.integers(25, 330))
        kind = int(rng.integers(0, 7))
        if kind == 0:
            s = rand_seq(L, rng.dirichlet([0.3] * 4))
        elif kind == 1:
            pat = pats[rng.integers(len(pats))]
            rep = (pat * (L // len(pat) + 1))[:int(L * rng.uniform(0.3, 0.8))]
            rest = rand_seq(L - len(rep))
            cut = int(rng.integers(0, len(rest) + 1))
            s = mutate(rest[:cut] + rep + rest[cut:], rng.choice([0, 0, 0.01, 0.03]))
        elif kind == 2:
            f = mutate(filt[rng.integers(38)], rng.choice([0, 0, 0.04]))
            body = rand_seq(max(L - 29, 1))
            s = f + body if rng.random() < 0.5 else body + f
        elif kind == 3:
            b, frac = "ACGT"[rng.integers(4)], rng.uniform(0.6, 0.8)
            s = "".join(b if rng.random() < frac else "ACGT"[rng.integers(4)] for _ in range(L))
        elif kind == 4:
            p1, p2, k = pats[rng.integers(len(pats))], pats[rng.integers(len(pats))], int(rng.integers(0, L))
            s = (p1 * L)[:k] + (p2 * L)[:L - k]
        elif kind == 5:
            s = rand_seq(L)
            if rng.random() < 0.5:
                s = s.lower()
            if rng.random() < 0.3:
                i = int(rng.integers(L))
                s = s[:i] + "N" + s[i + 1:]
        else:
            s = (pats[rng.integers(len(pats))] * L)[:L]
        cases.append(s)
    got = [host.test_read(s) for s in cases]
    want = [oracle.test_read(s.upper()) for s in cases]   # the reference upper-cases before testing (Dataset.cpp:303)
    assert 0.1 < sum(want) / len(want) < 0.9               # both outcomes well represented
    assert [s for s, a, b in zip(cases, got, want) if a != b] == []


def _write(path, text, gz=False):
    if gz:
        with gzip.open(path, "wt") as f:
            f.write(text)
    else:
        with open(path, "w") as f:
            f.write(text)


@pytest.mark.parametrize("gz", [False, True])
def test_fasta_fastq_parsing(tmp_path, gz):
    reads = synth.single_genome(50, 80, 10.0, seed=5).strings()
    recs = reads[:20] + ["ACGTNNNN" * 10] + reads[20:] + ["acgt" + reads[0][4:].lower()]
    ext = ".gz" if gz else ""
    fa = str(tmp_path / ("r.fa" + ext))
    # multi-line FASTA: sequence lines are joined (Dataset.cpp:276)
    _write(fa, "".join(f">r{i} desc\n{s[:30]}\n{s[30:]}\n" for i, s in enumerate(recs)), gz)
    fq = str(tmp_path / ("r.fq" + ext))
    _write(fq, "".join(f"@r{i}\n{s}\n+\n{'I' * len(s)}\n" for i, s in enumerate(recs)), gz)
    want, fi = oracle_filter(recs, 40)
    for path in (fa, fq):
        r = host.Reads(40)
        r.add_file(path)
        r.finalize()
        assert r.records == len(recs)
        assert list(r.file_index) == fi
        ref, _ = pack.pack_codes(synth.from_strings(want).codes, synth.from_strings(want).off, r.packed.shape[1])
        assert np.array_equal(r.packed, ref)
    # two files: file indices keep counting (Dataset.cpp:109-128)
    r = host.Reads(40)
    r.add_file(fa)
    r.add_file(fq)
    r.finalize()
    assert r.records == 2 * len(recs)
    assert list(r.file_index) == fi + [x + len(recs) for x in fi]
    with pytest.raises(host.HostError):
        host.Reads(40).add_file(str(tmp_path / "missing.fa"))


def test_pack_codes_matches_numpy():
    rs = synth.dup_contained(500, 150, 20.0, seed=3)
    a, la = host.pack_codes(rs.codes, rs.off)
    b, lb = pack.pack_codes(rs.codes, rs.off, a.shape[1])
    assert np.array_equal(a, b) and np.array_equal(la, lb)


def test_writers_match_reference_format(tmp_path):
    g = load_golden([p for p in GOLDEN if "fixture_contained_m30" in p][0])
    # edges / rows of the golden case, expressed in 0-based accepted-read ids
    reads, fi = oracle_filter(g["records"], 30)
    pos = {f: i for i, f in enumerate(fi)}
    lens = np.array([len(s) for s in reads], dtype=np.uint16)
    edges = np.zeros(len(g["ref_edges"]), dtype=gpu.EDGE_DTYPE)
    for k, line in enumerate(g["ref_edges"]):
        a, b, rest = line.split("\t")
        f = rest.split(",")
        edges[k] = (pos[int(a)], pos[int(b)], int(f[5]), int(f[0]))
    rows = np.zeros(len(g["ref_crows"]), dtype=gpu.CROW_DTYPE)
    for k, line in enumerate(g["ref_crows"]):
        a, b, rest = line.split("\t")
        f = rest.split(",")
        rows[k] = (pos[int(a)], pos[int(b)], int(f[0]), int(f[8]))
    pg, cr = str(tmp_path / "p_0_parGraph.txt"), str(tmp_path / "p_0_containedReads.txt")
    host.write_pargraph(pg, edges, np.array(fi), lens, flag=2)
    host.write_contained(cr, rows, np.array(fi), lens)
    assert [l.rstrip("\n") for l in open(pg)] == [l + ",2" for l in g["ref_edges"]]
    assert [l.rstrip("\n") for l in open(cr)] == g["ref_crows"]


def test_writers_many_lines_and_append(tmp_path):
    """the writers format blocks of lines on all cores: more lines than one round of blocks, order kept, append works"""
    rng = np.random.default_rng(5)
    n, ne = 50_000, 1_200_000
    fi = np.cumsum(rng.integers(1, 4, n)).astype(np.uint64)
    lens = rng.integers(60, 300, n).astype(np.uint16)
    e = np.zeros(ne, dtype=gpu.EDGE_DTYPE)
    e["src"] = rng.integers(0, n - 1, ne)
    e["dst"] = rng.integers(0, n, ne)
    e["offset"] = rng.integers(1, 59, ne)
    e["orient"] = rng.integers(0, 4, ne)
    path = str(tmp_path / "pg.txt")
    host.write_pargraph(path, e[:1_000_000], fi, lens, flag=2)
    host.write_pargraph(path, e[1_000_000:], fi, lens, flag=1, append=True)
    lines = open(path).read().split("\n")
    assert lines[-1] == "" and len(lines) == ne + 1
    for k in list(rng.integers(0, ne, 2000)) + [0, 65535, 65536, 999_999, 1_000_000, ne - 1]:
        s, d, off, o = (int(e[k][f]) for f in ("src", "dst", "offset", "orient"))
        sl, dl = int(lens[s]), int(lens[d])
        ovl = sl - off
        assert lines[k] == f"{fi[s]}\t{fi[d]}\t{o},{ovl},0,0,{sl},{off},{sl - 1},{dl},0,{ovl - 1},NA,{2 if k < 1_000_000 else 1}"
    r = np.zeros(200_000, dtype=gpu.CROW_DTYPE)
    r["contained"] = rng.integers(0, n, len(r)); r["container"] = rng.integers(0, n, len(r))
    r["orient"] = rng.integers(0, 4, len(r)); r["start"] = rng.integers(0, 50, len(r))
    cpath = str(tmp_path / "cr.txt")
    host.write_contained(cpath, r, fi, lens)
    cl = open(cpath).read().split("\n")
    assert len(cl) == len(r) + 1
    for k in list(rng.integers(0, len(r), 1000)) + [0, len(r) - 1]:
        a, b, o, st = (int(r[k][f]) for f in ("contained", "container", "orient", "start"))
        l2, l1 = int(lens[a]), int(lens[b])
        assert cl[k] == f"{fi[a]}\t{fi[b]}\t{o},{l2},0,0,{l2},0,{l2},{l1},{st},{st + l2}"
    e2 = e.copy()
    host.sort_edges(e2)
    assert np.array_equal(e2, gpu.sort_edges(e))


def test_sharded_pargraph_writer(tmp_path):
    """the reference's per-thread partial graphs: flags 2 / 0 / 1 by read range, every edge of a node in its shard's file"""
    from helpers import check_partial_graphs
    rng = np.random.default_rng(8)
    n, ne = 20_000, 150_000
    fi = np.cumsum(rng.integers(1, 3, n)).astype(np.uint64)
    lens = rng.integers(60, 300, n).astype(np.uint16)
    src = rng.integers(0, n - 1, ne)
    dst = np.minimum(src + 1 + rng.geometric(0.002, ne), n - 1)     # mostly near neighbours, some far ones
    pairs = np.unique(np.stack([src, dst], axis=1)[src < dst], axis=0)
    e = np.zeros(len(pairs), dtype=gpu.EDGE_DTYPE)
    e["src"], e["dst"] = pairs[:, 0], pairs[:, 1]
    e["offset"] = rng.integers(1, 59, len(e)); e["orient"] = rng.integers(0, 4, len(e))
    e = gpu.sort_edges(e)
    want = None
    for shards in (1, 2, 3, 16):
        prefix = str(tmp_path / f"s{shards}")
        host.write_pargraph_sharded(prefix, shards, e, n, fi, lens)
        got = check_partial_graphs(prefix, shards)
        if want is None:
            one = str(tmp_path / "one.txt")
            host.write_pargraph(one, e, fi, lens, flag=2)
            want = set(l.rstrip("\n").rsplit(",", 1)[0] for l in open(one))
            assert [l.rstrip("\n") for l in open(prefix + "_0_parGraph.txt")] == [l.rstrip("\n") for l in open(one)]
        assert got == want and len(got) == len(e)
        for t in range(shards):          # shard t owns a contiguous range of reads
            lo, hi = n * t // shards, n * (t + 1) // shards
            for line in list(open(f"{prefix}_{t}_parGraph.txt"))[:2000]:
                a, b, rest = line.rstrip("\n").split("\t")
                flag = int(rest.rsplit(",", 1)[1])
                sa, sb = int(np.searchsorted(fi, int(a))), int(np.searchsorted(fi, int(b)))
                assert (lo <= sa < hi) if flag in (0, 2) else not (lo <= sa < hi)
                assert (lo <= sb < hi) if flag in (1, 2) else not (lo <= sb < hi)
    # unsorted input is refused, an empty edge list gives empty files
    import pytest as _pt
    with _pt.raises(host.HostError):
        host.write_pargraph_sharded(str(tmp_path / "bad"), 2, e[::-1].copy(), n, fi, lens)
    host.write_pargraph_sharded(str(tmp_path / "empty"), 3, e[:0], n, fi, lens)
    assert all(open(str(tmp_path / f"empty_{t}_parGraph.txt")).read() == "" for t in range(3))
