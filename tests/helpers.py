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
