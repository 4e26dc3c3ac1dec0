"""Downstream acceptance: the reference's own parsimplify (oracle/_ref/parsimplify) must accept the files
we write and contract them to the same graph it gets from the reference's own parGraph file (committed in
tests/golden/*.npz as ref_parsimplify): as one file with mark flag 2, and as the per-thread partial graphs `buildG -t n`
writes (mark flags 2 / 0 / 1; one parsimplify per file)."""
import os
import subprocess
import numpy as np
import pytest
from helpers import GOLDEN, load_golden, oracle_forms, HERE
from disco_b200 import gpu, host

PARSIMPLIFY = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "parsimplify")
CASES = [p for p in GOLDEN if "capfire" not in p]


def _canon(lines):
    """A composite edge lists the reads it swallowed in traversal order; compare order-free."""
    out = []
    for l in lines:
        f = l.split("\t")
        tail = sorted(f[-1].strip("()").split(")(")) if len(f) > 3 and f[-1].startswith("(") else f[-1:]
        out.append("\t".join(f[:-1]) + "\t" + ",".join(tail))
    return sorted(out)


def _run_parsimplify(tmp_path, edges, fi, lens, m):
    pg = str(tmp_path / "o_0_parGraph.txt")
    host.write_pargraph(pg, host.sort_edges(edges), np.asarray(fi, dtype=np.uint64), np.asarray(lens, dtype=np.uint16), flag=2)
    out = str(tmp_path / "simple.txt")
    r = subprocess.run([PARSIMPLIFY, pg, out, str(m), "1"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return open(out).read().splitlines()


@pytest.mark.skipif(not os.access(PARSIMPLIFY, os.X_OK), reason="oracle/_ref/parsimplify not built")
@pytest.mark.parametrize("path", CASES, ids=lambda p: p.split("/")[-1][:-4])
def test_writer_output_is_accepted_cpu(tmp_path, path):
    """CPU: oracle edges through OUR writer."""
    g = load_golden(path)
    if not g["ref_parsimplify"]:
        pytest.skip("golden has no parsimplify output")
    o = oracle_forms(g["records"], g["min_overlap"])
    e = np.zeros(len(o["res"].edges), dtype=gpu.EDGE_DTYPE)
    e["src"] = o["res"].edges["src"] - 1
    e["dst"] = o["res"].edges["dst"] - 1
    e["offset"] = o["res"].edges["offset"]
    e["orient"] = o["res"].edges["orient"]
    got = _run_parsimplify(tmp_path, e, o["fi"], o["lens"], g["min_overlap"])
    assert _canon(got) == _canon(g["ref_parsimplify"])


@pytest.mark.gpu
@pytest.mark.skipif(not os.access(PARSIMPLIFY, os.X_OK), reason="oracle/_ref/parsimplify not built")
@pytest.mark.parametrize("path", CASES[:4], ids=lambda p: p.split("/")[-1][:-4])
def test_gpu_files_are_accepted(tmp_path, path):
    from disco_b200.buildgraph import BuildGraph
    g = load_golden(path)
    if not g["ref_parsimplify"]:
        pytest.skip("golden has no parsimplify output")
    bg = BuildGraph(min_overlap=g["min_overlap"])
    bg.add_records(g["records"])
    res = bg.run()
    try:
        got = _run_parsimplify(tmp_path, res.edges, res.file_index, res.lens, g["min_overlap"])
        assert _canon(got) == _canon(g["ref_parsimplify"])
    finally:
        bg.close()


@pytest.mark.skipif(not os.access(PARSIMPLIFY, os.X_OK), reason="oracle/_ref/parsimplify not built")
@pytest.mark.parametrize("path", [p for p in CASES if "single" in p or "paired" in p][:2], ids=lambda p: p.split("/")[-1][:-4])
def test_partial_graphs_are_accepted_cpu(tmp_path, path):
    """CPU: oracle edges through the sharded writer; the reference parsimplify runs on every partial graph.  Nodes marked
    in a file are contracted there, so no read may be swallowed by composite edges of two different files, and a read
    swallowed in a file is never an endpoint of an edge that file writes."""
    from helpers import check_partial_graphs
    g = load_golden(path)
    o = oracle_forms(g["records"], g["min_overlap"])
    e = np.zeros(len(o["res"].edges), dtype=gpu.EDGE_DTYPE)
    e["src"] = o["res"].edges["src"] - 1
    e["dst"] = o["res"].edges["dst"] - 1
    e["offset"] = o["res"].edges["offset"]
    e["orient"] = o["res"].edges["orient"]
    e = gpu.sort_edges(e)
    shards = 3
    prefix = str(tmp_path / "p")
    host.write_pargraph_sharded(prefix, shards, e, len(o["lens"]), np.asarray(o["fi"], dtype=np.uint64), np.asarray(o["lens"], dtype=np.uint16))
    assert check_partial_graphs(prefix, shards) == set(o["edges"])
    swallowed = []
    for t in range(shards):
        out = str(tmp_path / f"simple{t}.txt")
        r = subprocess.run([PARSIMPLIFY, f"{prefix}_{t}_parGraph.txt", out, str(g["min_overlap"]), "1"], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        inner, ends = set(), set()
        for l in open(out).read().splitlines():
            f = l.split("\t")
            ends.update((int(f[0]), int(f[1])))
            if len(f) > 3 and f[-1].startswith("("):
                inner.update(int(x.split(",")[0]) for x in f[-1].strip("()").split(")("))
        assert not (inner & ends)
        swallowed.append(inner)
    assert sum(len(x) for x in swallowed) > 0
    for a in range(shards):
        for b in range(a + 1, shards):
            assert not (swallowed[a] & swallowed[b])
