"""Regenerates tests/golden/*.npz.  Runs only where /root/reference exists (the build container):
each case = raw input records + the output of the REAL reference binary (oracle/_ref/buildG, -t 1) on them.

    python tests/golden/make_golden.py
"""
import os
import sys
import tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402
from disco_b200 import synth  # noqa: E402

REF_FIX = "/root/reference/src/BuildGraph"


def parse_fasta(path):
    recs = []
    for blk in open(path).read().split(">")[1:]:
        lines = blk.split("\n")
        recs.append("".join(lines[1:]).upper())
    return recs


def make(name, records, min_overlap, paired=False):
    d = tempfile.mkdtemp(prefix="golden_")
    fa = os.path.join(d, "r.fa")
    with open(fa, "w") as f:
        for i, s in enumerate(records):
            f.write(f">{i + 1}\n{s}\n")
    ref = oracle.run_ref([fa], os.path.join(d, "o"), min_overlap, threads=1, paired=paired)
    assert "Graph construction complete" in ref["log"], ref["log"][-2000:]
    # downstream acceptance: what the reference's own parsimplify makes of the reference's own file
    simple = []
    ps = os.path.join(ROOT, "oracle", "_ref", "parsimplify")
    if os.access(ps, os.X_OK):
        import subprocess
        out = os.path.join(d, "simple.txt")
        r = subprocess.run([ps, os.path.join(d, "o_0_parGraph.txt"), out, str(min_overlap), "1"], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        simple = sorted(open(out).read().splitlines())
    np.savez_compressed(os.path.join(HERE, name + ".npz"),
                        records=np.array(records, dtype=object), min_overlap=min_overlap,
                        ref_edges=np.array(ref["edges"], dtype=object),
                        ref_crows=np.array(ref["contained_rows"], dtype=object),
                        ref_parsimplify=np.array(simple, dtype=object))
    print(f"{name}: {len(records)} records, m={min_overlap}: {len(ref['edges'])} edges, {len(ref['contained_rows'])} contained rows")


def main():
    oracle.build()
    make("fixture_contained_m30", parse_fasta(f"{REF_FIX}/10reads_containedReads.fasta"), 30)
    make("fixture_forward_m30", parse_fasta(f"{REF_FIX}/10reads_forward.fasta"), 30)
    make("single_3000x150_m50", synth.single_genome(3000, 150, 30.0, seed=1).strings(), 50)
    dup = synth.dup_contained(3000, 150, 60.0, seed=4).strings()
    for m in (35, 50, 75):
        make(f"dup_contained_3000_m{m}", dup, m)
    make("paired_800x2x250_m30", synth.paired_genome(800, 250, seed=1).strings(), 30, paired=True)
    make("metagenome_4000_m50", synth.metagenome(4000, 8, 6000, 150, seed=3).strings(), 50)
    make("repeats_capfire_2000_m50", synth.repeats(2000, 150, seed=11).strings(), 50)
    # records the reference filter rejects (N, homopolymer, micro-repeat ends, too short) mixed into good reads
    rng = np.random.default_rng(9)
    good = synth.single_genome(600, 120, 25.0, seed=9).strings()
    bad = ["ACGTN" * 24, "A" * 100 + "ACGT" * 5, "ACACACACACACACACACACACACACACA" + good[0][29:], good[1][:40],
           "AT" * 60, good[2][:-29] + "TTCTTCTTCTTCTTCTTCTTCTTCTTCTT", "acgt" * 30, good[3].lower()]
    recs = list(good)
    for b in bad:
        recs.insert(int(rng.integers(0, len(recs))), b)
    make("filter_mix_m40", recs, 40)
    # variable lengths incl. > 512 bp (generic long-read matcher) and an even K
    g = synth.random_genome(rng, 20000)
    recs = []
    for _ in range(900):
        L = int(rng.integers(60, 700))
        s = int(rng.integers(0, len(g) - L))
        r = g[s:s + L]
        if rng.random() < 0.5:
            r = synth.revcomp_codes(r)
        recs.append("".join("ACGT"[c] for c in r))
    make("varlen_60_700_m35", recs, 35)


if __name__ == "__main__":
    main()
