"""ctypes wrapper around oracle/disco_oracle.c and the real reference binary (oracle/_ref/buildG).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package (disco_b200) never imports this module.
"""
import ctypes as C
import os
import re
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libdisco_oracle.so")
REF_BIN = os.path.join(HERE, "_ref", "buildG")


def build(quiet=True):
    subprocess.run(["make", "-s", "-C", HERE, "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None, stderr=subprocess.DEVNULL if quiet else None)


class _CRow(C.Structure):
    _fields_ = [("contained", C.c_uint64), ("container", C.c_uint64), ("orient", C.c_uint32),
                ("len2", C.c_uint32), ("len1", C.c_uint32), ("start", C.c_uint32)]


class _Edge(C.Structure):
    _fields_ = [("src", C.c_uint64), ("dst", C.c_uint64), ("orient", C.c_uint32), ("offset", C.c_uint32)]


class _Stats(C.Structure):
    _fields_ = [("cap_fired", C.c_uint64), ("multi_overlap_pairs", C.c_uint64), ("one_sided_edges", C.c_uint64),
                ("raw_directed", C.c_uint64), ("lookups", C.c_uint64), ("candidates", C.c_uint64)]


CROW_DTYPE = np.dtype([("contained", "<u8"), ("container", "<u8"), ("orient", "<u4"), ("len2", "<u4"),
                       ("len1", "<u4"), ("start", "<u4")])
EDGE_DTYPE = np.dtype([("src", "<u8"), ("dst", "<u8"), ("orient", "<u4"), ("offset", "<u4")])

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.oracle_test_read.argtypes = [C.c_char_p, C.c_uint64]
        L.oracle_test_read.restype = C.c_int
        L.oracle_create.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32]
        L.oracle_create.restype = C.c_void_p
        L.oracle_free.argtypes = [C.c_void_p]
        L.oracle_contained.argtypes = [C.c_void_p, C.POINTER(C.POINTER(_CRow))]
        L.oracle_contained.restype = C.c_uint64
        L.oracle_super_read.argtypes = [C.c_void_p]
        L.oracle_super_read.restype = C.POINTER(C.c_uint64)
        L.oracle_raw_edges.argtypes = [C.c_void_p, C.POINTER(C.POINTER(_Edge)), C.POINTER(_Stats)]
        L.oracle_raw_edges.restype = C.c_uint64
        L.oracle_reduce.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.POINTER(_Edge)), C.POINTER(_Stats)]
        L.oracle_reduce.restype = C.c_uint64
        L.oracle_free_buf.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def test_read(seq: str) -> bool:
    b = seq.encode()
    return bool(lib().oracle_test_read(b, len(b)))


def _take(ptr, n, dtype):
    if n == 0:
        out = np.zeros(0, dtype=dtype)
    else:
        buf = (C.c_char * (n * dtype.itemsize)).from_address(C.addressof(ptr.contents))
        out = np.frombuffer(buf, dtype=dtype).copy()
    lib().oracle_free_buf(ptr)
    return out


class OracleResult:
    """super_read[N] (0 / 1-based container), crows, raw (directed finds), edges (reduced, src<dst), stats dict."""


def run(reads, min_overlap: int, reduce: bool = True) -> OracleResult:
    """reads: list of upper-case ACGT strings that already passed the read filter (ids 1..N in list order)."""
    L = lib()
    lens = np.array([len(r) for r in reads], dtype=np.uint64)
    off = np.zeros(len(reads) + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    bases = "".join(reads).encode()
    buf = C.create_string_buffer(bases, len(bases) + 1)
    ctx = L.oracle_create(C.cast(buf, C.c_void_p), off.ctypes.data_as(C.c_void_p), len(reads), min_overlap)
    res = OracleResult()
    try:
        rows = C.POINTER(_CRow)()
        n = L.oracle_contained(ctx, C.byref(rows))
        res.crows = _take(rows, n, CROW_DTYPE)
        sp = L.oracle_super_read(ctx)
        res.super_read = np.ctypeslib.as_array(sp, shape=(max(len(reads), 1),))[:len(reads)].copy()
        st = _Stats()
        ep = C.POINTER(_Edge)()
        n = L.oracle_raw_edges(ctx, C.byref(ep), C.byref(st))
        raw_addr = C.addressof(ep.contents) if n else None
        if reduce:
            kp = C.POINTER(_Edge)()
            nk = L.oracle_reduce(ctx, raw_addr, n, C.byref(kp), C.byref(st))
            res.edges = _take(kp, nk, EDGE_DTYPE)
        else:
            res.edges = None
        res.raw = _take(ep, n, EDGE_DTYPE)
        res.stats = {f: getattr(st, f) for f, _ in _Stats._fields_}
    finally:
        L.oracle_free(ctx)
    return res


# ------------------------------------------------------------------ text forms
def edge_lines(edges, file_index, lens):
    """canonical parGraph lines without the trailing mark flag; edges: structured array src<dst (1-based ids)."""
    out = []
    for e in edges:
        s, d = int(e["src"]), int(e["dst"])
        sl, dl, off = int(lens[s - 1]), int(lens[d - 1]), int(e["offset"])
        ovl = sl - off
        out.append(f"{int(file_index[s - 1])}\t{int(file_index[d - 1])}\t{int(e['orient'])},{ovl},0,0,{sl},{off},{sl - 1},{dl},0,{ovl - 1},NA")
    return sorted(out)


def crow_lines(crows, file_index):
    return [f"{int(file_index[int(r['contained']) - 1])}\t{int(file_index[int(r['container']) - 1])}\t"
            f"{int(r['orient'])},{int(r['len2'])},0,0,{int(r['len2'])},0,{int(r['len2'])},{int(r['len1'])},"
            f"{int(r['start'])},{int(r['start']) + int(r['len2'])}" for r in crows]


# ------------------------------------------------------------------ the real reference binary
def have_ref() -> bool:
    return os.access(REF_BIN, os.X_OK)


def run_ref(fasta_files, out_prefix, min_overlap, threads=1, paired=False, mem_gb=8, timeout=3600):
    """Runs oracle/_ref/buildG; returns dict(edges=sorted canonical lines w/o flag, contained_rows=[...] in file
    order of shard 0.., contained_set=set of file indices, times={function: seconds}, log=str)."""
    os.makedirs(os.path.dirname(os.path.abspath(out_prefix)), exist_ok=True)
    cfg = out_prefix + "_oracle.cfg"
    with open(cfg, "w") as f:
        f.write(f"MinOverlap4BuildGraph = {min_overlap}\n")
    cmd = [REF_BIN, "-pe" if paired else "-se", ",".join(fasta_files), "-f", out_prefix, "-p", cfg,
           "-t", str(threads), "-m", str(mem_gb)]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
    log = p.stdout + p.stderr
    edges = set()
    rows = []
    for t in range(threads):
        pg = f"{out_prefix}_{t}_parGraph.txt"
        if os.path.exists(pg):
            with open(pg) as f:
                for line in f:
                    line = line.rstrip("\n")
                    if line:
                        edges.add(re.sub(r",[012]$", "", line))
        cr = f"{out_prefix}_{t}_containedReads.txt"
        if os.path.exists(cr):
            with open(cr) as f:
                rows += [l.rstrip("\n") for l in f if l.strip()]
    times = {m.group(1): float(m.group(2)) for m in re.finditer(r"Function (\w+)\(\) finished in ([0-9.eE+-]+) Seconds", log)}
    return dict(edges=sorted(edges), contained_rows=rows,
                contained_set=set(int(r.split("\t")[0]) for r in rows), times=times, log=log, returncode=p.returncode)
