"""parsimplify's composite-edge contraction + dead-end removal on the GPU (disco_b200/csrc/simplify.cu), from
the reduced edges still in HBM, against the REAL reference parsimplify (oracle/_ref/parsimplify, one thread) run on the
parGraph file the same run wrote -- line for line."""
import os
import subprocess
import numpy as np
import pytest
from helpers import GOLDEN, load_golden, HERE
from disco_b200 import host, synth
from disco_b200.buildgraph import BuildGraph

pytestmark = pytest.mark.gpu
PARSIMPLIFY = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "parsimplify")
needs_ref = pytest.mark.skipif(not os.access(PARSIMPLIFY, os.X_OK), reason="oracle/_ref/parsimplify not built")


def _reference_lines(tmp_path, res, m):
    pg = str(tmp_path / "o_0_parGraph.txt")
    host.write_pargraph(pg, host.sort_edges(res.edges), res.file_index, res.lens, flag=2)
    out = str(tmp_path / "simple.txt")
    r = subprocess.run([PARSIMPLIFY, pg, out, str(m), "1"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return [l.rstrip("\n") for l in open(out)]


def _check(records, m, tmp_path, expect_composite=True):
    bg = BuildGraph(min_overlap=m)
    bg.add_records(records)
    res = bg.run()
    try:
        mine = bg.simplified_lines(min_overlap=m)
        st = bg.simplify_stats
        ref = _reference_lines(tmp_path, res, m)
        if st["cycle_edges"] == 0:
            # line for line: ends, orientation, offset sum, edge length, every inner read in order (the reference binary is
            # built with its copyEdge determinism fix, oracle/build_ref.sh patch 3)
            assert sorted(mine) == sorted(ref)
        else:   # an isolated cycle: the reference breaks it where its node order starts (simplify.cu header)
            assert len(mine) >= len(ref)
        if expect_composite:
            assert any(l.split("\t")[3] for l in mine)
        return st, mine
    finally:
        bg.close()


@needs_ref
@pytest.mark.parametrize("path", [p for p in GOLDEN if "capfire" not in p], ids=lambda p: p.split("/")[-1][:-4])
def test_goldens(tmp_path, path):
    g = load_golden(path)
    _check(g["records"], g["min_overlap"], tmp_path, expect_composite=False)


@needs_ref
@pytest.mark.parametrize("name,make,m", [
    ("single_20k", lambda: synth.single_genome(20000, 150, 30.0, seed=21), 50),
    ("single_low_coverage", lambda: synth.single_genome(6000, 150, 8.0, seed=22), 50),    # many contig ends and short tips
    ("metagenome", lambda: synth.metagenome(30000, 12, 30000, 150, seed=26), 50),
    ("varlen_dups", lambda: synth.dup_contained(12000, 150, 40.0, seed=23), 35),
    ("paired250", lambda: synth.paired_genome(4000, 250, seed=25), 30),
    ("repeats", lambda: synth.repeats(4000, 150, seed=31), 50),                           # branching nodes, dead ends
], ids=lambda x: x if isinstance(x, str) else None)
def test_synthetic(tmp_path, name, make, m):
    st, mine = _check(make().strings(), m, tmp_path)
    assert st["rounds"] >= 1


@needs_ref
def test_one_genome_is_one_composite_edge(tmp_path):
    """SURVEY 8b sanity check: a single random genome at 30x contracts to (about) one edge about as long as the genome."""
    rs = synth.single_genome(40000, 150, 30.0, seed=5)
    st, mine = _check(rs.strings(), 50, tmp_path)
    longest = max(int(l.split("\t")[2].split(",")[2]) for l in mine)
    assert longest > 0.9 * 40000 * 150 / 30.0
