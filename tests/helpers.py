"""Shared test helpers: golden loading, oracle-vs-product comparison forms."""
import glob
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = sorted(glob.glob(os.path.join(HERE, "golden", "*.npz")))


def load_golden(path):
    z = np.load(path, allow_pickle=True)
    return dict(name=os.path.basename(path)[:-4], records=[str(x) for x in z["records"]], min_overlap=int(z["min_overlap"]),
                ref_edges=[str(x) for x in z["ref_edges"]], ref_crows=[str(x) for x in z["ref_crows"]],
                ref_parsimplify=[str(x) for x in z["ref_parsimplify"]] if "ref_parsimplify" in z.files else None)


def oracle_filter(records, min_overlap):
    """(accepted upper-case reads, their 1-based file indices) using the ORACLE's filter."""
    from oracle import oracle
    reads, fi = [], []
    for i, s in enumerate(records):
        u = s.upper()
        if len(u) > min_overlap and oracle.test_read(u):
            reads.append(u)
            fi.append(i + 1)
    return reads, fi


def oracle_forms(records, min_overlap):
    from oracle import oracle
    reads, fi = oracle_filter(records, min_overlap)
    res = oracle.run(reads, min_overlap)
    lens = [len(r) for r in reads]
    return dict(res=res, reads=reads, fi=fi, lens=lens, edges=oracle.edge_lines(res.edges, fi, lens),
                crows=oracle.crow_lines(res.crows, fi),
                contained=set(int(fi[i]) for i in np.nonzero(res.super_read)[0]))


def check_partial_graphs(prefix, shards):
    """The invariants parsimplify relies on when it gets one parGraph file per BuildGraph thread (OverlapGraph.cpp:826-859,
    OverlapGraphSimple.cpp:632-641): a node is marked in exactly one file, and every edge of a node marked in a file is in
    that file; an edge whose endpoints are marked in one file is there once with flag 2, any other edge twice -- flag 0 in
    the source's file, flag 1 in the destination's.  Returns the set of canonical lines (flag stripped)."""
    files = []
    for t in range(shards):
        rows = []
        with open(f"{prefix}_{t}_parGraph.txt") as f:
            for line in f:
                body, flag = line.rstrip("\n").rsplit(",", 1)
                src, dst = body.split("\t")[:2]
                rows.append((body, int(src), int(dst), int(flag)))
        files.append(rows)
    marked = [set() for _ in range(shards)]
    for t, rows in enumerate(files):
        for _, s, d, fl in rows:
            assert fl in (0, 1, 2)
            if fl in (0, 2):
                marked[t].add(s)
            if fl in (1, 2):
                marked[t].add(d)
    for a in range(shards):
        for b in range(a + 1, shards):
            assert not (marked[a] & marked[b]), "a node is marked in two files"
    home = {v: t for t in range(shards) for v in marked[t]}
    seen = {}
    for t, rows in enumerate(files):
        for body, s, d, fl in rows:
            assert body not in seen.get(t, set()), "duplicate line in one file"
            seen.setdefault(t, set()).add(body)
            if fl == 2:
                assert home[s] == t and home[d] == t
            elif fl == 0:
                assert home[s] == t and home[d] != t
            else:
                assert home[d] == t and home[s] != t
    allb = {}
    for t, rows in enumerate(files):
        for body, s, d, fl in rows:
            allb.setdefault(body, []).append((t, fl))
    for body, where in allb.items():
        flags = sorted(fl for _, fl in where)
        assert flags in ([2], [0, 1]), (body, where)
    return set(allb)
