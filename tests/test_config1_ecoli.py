"""BASELINE config 1 stand-in (the reference's test/Ecoli_250_500_test.fna is not in the mount): seeded
random genome, interleaved 2 x 250 bp pairs (-pe), insert 500 +- 50, 30x.  Our buildG executable against the REAL
reference binary (oracle/_ref/buildG, all host cores) run on the same FASTA in the same test: canonical edge set,
contained set, and -- per file -- the line formats."""
import os
import re
import subprocess
import pytest
from helpers import HERE
from disco_b200 import synth

ROOT = os.path.dirname(HERE)
BUILDG = os.path.join(ROOT, "disco_b200", "bin", "buildG")
REF = os.path.join(ROOT, "oracle", "_ref", "buildG")
pytestmark = pytest.mark.gpu


def _canon_edges(prefix, shards):
    out = set()
    for t in range(shards):
        with open(f"{prefix}_{t}_parGraph.txt") as f:
            for line in f:
                out.add(re.sub(r",[012]$", "", line.rstrip("\n")))
    return out


def _contained(prefix, shards):
    out = {}
    for t in range(shards):
        with open(f"{prefix}_{t}_containedReads.txt") as f:
            for line in f:
                a = line.split("\t")
                out[int(a[0])] = line.rstrip("\n")
    return out


@pytest.mark.skipif(not os.access(REF, os.X_OK), reason="oracle/_ref/buildG not built")
@pytest.mark.parametrize("n_pairs,min_overlap", [(60_000, 30), (60_000, 50), (276_000, 30)],
                         ids=["1Mb_m30", "1Mb_m50", "ecoli_size_4.6Mb_552k_reads_m30"])
def test_paired_genome_against_reference_binary(tmp_path, n_pairs, min_overlap):
    # 4.6 Mb genome, 2 x 250 bp pairs, 30x = 552 k reads, minOverlap 30 (disco.cfg:9), runEcoli.sh's -n 4 -m 5
    rs = synth.paired_genome(n_pairs, 250, insert=500, insert_sd=50, coverage=30.0, seed=1)
    fa = str(tmp_path / "ecoli_like.fna")
    rs.write_fasta(fa)
    cfg = tmp_path / "disco.cfg"
    cfg.write_text(f"MinOverlap4BuildGraph = {min_overlap}\n")
    cores = os.cpu_count() or 1
    os.makedirs(tmp_path / "ours")
    os.makedirs(tmp_path / "ref")
    r1 = subprocess.run([BUILDG, "-pe", fa, "-f", str(tmp_path / "ours" / "g"), "-p", str(cfg), "-t", "4", "-m", "5"],
                        capture_output=True, text=True)
    assert r1.returncode == 0, r1.stdout[-2000:] + r1.stderr[-2000:]
    r2 = subprocess.run([REF, "-pe", fa, "-f", str(tmp_path / "ref" / "g"), "-p", str(cfg), "-t", str(cores), "-m", "64"],
                        capture_output=True, text=True)
    assert "Graph construction complete" in r2.stdout
    m = re.search(r"cap_fired (\d+), multi_overlap_pairs (\d+), one_sided_edges (\d+)", r1.stdout)
    assert m and all(int(x) == 0 for x in m.groups()), r1.stdout[-800:]
    ours, ref = _canon_edges(str(tmp_path / "ours" / "g"), 4), _canon_edges(str(tmp_path / "ref" / "g"), cores)
    assert ours == ref                                        # bit-exact canonical edge set (src, dst, orientation, offsets)
    oc, rc = _contained(str(tmp_path / "ours" / "g"), 4), _contained(str(tmp_path / "ref" / "g"), cores)
    assert set(oc) == set(rc)                                 # same contained-read set (rows are racy in the reference beyond -t 1)
    assert open(str(tmp_path / "ours" / "g") + "_ReadIDMap.txt").read() == open(str(tmp_path / "ref" / "g") + "_ReadIDMap.txt").read()
