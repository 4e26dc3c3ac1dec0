"""ctypes binding of include/disco_host.h (libdisco_host.so): read filter, parsing, packing, writers."""
import ctypes as C
import os
import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libdisco_host.so")

EXPORTS = ["disco_host_last_error", "disco_host_test_read", "disco_reads_new", "disco_reads_free", "disco_reads_add_file",
           "disco_reads_add_records", "disco_reads_finalize", "disco_reads_count", "disco_reads_records",
           "disco_reads_words_per_read", "disco_reads_packed", "disco_reads_len", "disco_reads_file_index",
           "disco_reads_min_len", "disco_reads_max_len", "disco_host_pack_codes", "disco_write_pargraph",
           "disco_write_contained", "disco_host_sort_contained", "disco_host_sort_edges", "disco_write_pargraph_sharded"]

_lib = None


class HostError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise HostError(f"{LIB_PATH} not built: run `python -m disco_b200.build`")
        L = C.CDLL(LIB_PATH)
        vp, u64, u32, i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int
        L.disco_host_last_error.restype = C.c_char_p
        L.disco_host_test_read.argtypes = [C.c_char_p, u64]
        L.disco_reads_new.argtypes = [u32, i32]
        L.disco_reads_new.restype = vp
        L.disco_reads_free.argtypes = [vp]
        L.disco_reads_free.restype = None
        L.disco_reads_add_file.argtypes = [vp, C.c_char_p]
        L.disco_reads_add_records.argtypes = [vp, vp, vp, u64]
        L.disco_reads_finalize.argtypes = [vp]
        for f in ("disco_reads_count", "disco_reads_records"):
            getattr(L, f).argtypes = [vp]
            getattr(L, f).restype = u64
        for f in ("disco_reads_words_per_read", "disco_reads_min_len", "disco_reads_max_len"):
            getattr(L, f).argtypes = [vp]
            getattr(L, f).restype = u32
        for f in ("disco_reads_packed", "disco_reads_len", "disco_reads_file_index"):
            getattr(L, f).argtypes = [vp]
            getattr(L, f).restype = vp
        L.disco_host_pack_codes.argtypes = [vp, vp, u64, u32, vp, vp, i32]
        L.disco_write_pargraph.argtypes = [C.c_char_p, vp, u64, vp, vp, i32, i32]
        L.disco_write_contained.argtypes = [C.c_char_p, vp, u64, vp, vp, i32]
        L.disco_write_pargraph_sharded.argtypes = [C.c_char_p, u32, vp, u64, u64, vp, vp]
        L.disco_host_sort_contained.argtypes = [vp, u64, vp, u32]
        L.disco_host_sort_edges.argtypes = [vp, u64]
        _lib = L
    return _lib


def _ck(rc):
    if rc:
        raise HostError(lib().disco_host_last_error().decode())


def test_read(seq: str) -> bool:
    b = seq.encode()
    return bool(lib().disco_host_test_read(b, len(b)))


class Reads:
    """Accepted reads of a data set in file order (= the reference's Dataset + the packed payload of its HashTable)."""

    def __init__(self, min_overlap: int, threads: int = 0):
        self._L = lib()
        self._h = self._L.disco_reads_new(min_overlap, threads)
        self.min_overlap = min_overlap
        self._final = False

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.disco_reads_free(self._h)
            self._h = None

    def add_file(self, path: str):
        _ck(self._L.disco_reads_add_file(self._h, path.encode()))

    def add_records(self, records):
        """records: list of raw sequence strings (any case, may contain non-ACGT)."""
        lens = np.array([len(r) for r in records], dtype=np.uint64)
        off = np.zeros(len(records) + 1, dtype=np.uint64)
        np.cumsum(lens, out=off[1:])
        blob = "".join(records).encode()
        _ck(self._L.disco_reads_add_records(self._h, blob, off.ctypes.data, len(records)))

    def finalize(self):
        _ck(self._L.disco_reads_finalize(self._h))
        self._final = True
        return self

    @property
    def n(self):
        return self._L.disco_reads_count(self._h)

    @property
    def records(self):
        return self._L.disco_reads_records(self._h)

    def _arr(self, ptr, dtype, shape):
        n = int(np.prod(shape))
        if n == 0:
            return np.zeros(shape, dtype=dtype)
        buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
        return np.frombuffer(buf, dtype=dtype).reshape(shape)

    @property
    def packed(self):
        assert self._final
        w = self._L.disco_reads_words_per_read(self._h)
        return self._arr(self._L.disco_reads_packed(self._h), np.uint64, (self.n, w))

    @property
    def lens(self):
        assert self._final
        return self._arr(self._L.disco_reads_len(self._h), np.uint16, (self.n,))

    @property
    def file_index(self):
        return self._arr(self._L.disco_reads_file_index(self._h), np.uint64, (self.n,))


def pack_codes(codes: np.ndarray, off: np.ndarray, words_per_read: int = None, out=None, lens_out=None, threads: int = 0):
    """Multi-threaded packer: codes uint8 (0..3) concatenated, off n+1 offsets -> (uint64[n,wpr], uint16[n])."""
    off = np.ascontiguousarray(off, dtype=np.uint64)
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    n = len(off) - 1
    if words_per_read is None:
        mx = int(np.diff(off.astype(np.int64)).max()) if n else 1
        words_per_read = max(1, (mx + 31) // 32)
    if out is None:
        out = np.empty((n, words_per_read), dtype=np.uint64)
    if lens_out is None:
        lens_out = np.empty(n, dtype=np.uint16)
    _ck(lib().disco_host_pack_codes(codes.ctypes.data, off.ctypes.data, n, words_per_read, out.ctypes.data,
                                    lens_out.ctypes.data, threads))
    return out, lens_out


def write_pargraph(path, edges, file_index, lens, flag=2, append=False):
    edges = np.ascontiguousarray(edges)
    fi = np.ascontiguousarray(file_index, dtype=np.uint64)
    ln = np.ascontiguousarray(lens, dtype=np.uint16)
    _ck(lib().disco_write_pargraph(path.encode(), edges.ctypes.data, len(edges), fi.ctypes.data, ln.ctypes.data, flag, int(append)))


def write_pargraph_sharded(prefix, shards, edges, n_reads, file_index, lens):
    """<prefix>_<t>_parGraph.txt, t = 0..shards-1: the reference's per-thread partial graphs with their mark flags
    (edges sorted by (src, dst))"""
    edges = np.ascontiguousarray(edges)
    fi = np.ascontiguousarray(file_index, dtype=np.uint64)
    ln = np.ascontiguousarray(lens, dtype=np.uint16)
    _ck(lib().disco_write_pargraph_sharded(prefix.encode(), shards, edges.ctypes.data, len(edges), n_reads, fi.ctypes.data, ln.ctypes.data))


def write_contained(path, rows, file_index, lens, append=False):
    rows = np.ascontiguousarray(rows)
    fi = np.ascontiguousarray(file_index, dtype=np.uint64)
    ln = np.ascontiguousarray(lens, dtype=np.uint16)
    _ck(lib().disco_write_contained(path.encode(), rows.ctypes.data, len(rows), fi.ctypes.data, ln.ctypes.data, int(append)))


def sort_contained(rows: np.ndarray, lens: np.ndarray, min_overlap: int) -> np.ndarray:
    """In place: the reference's -t 1 emission order (container, position, record)."""
    rows = np.ascontiguousarray(rows)
    ln = np.ascontiguousarray(lens, dtype=np.uint16)
    _ck(lib().disco_host_sort_contained(rows.ctypes.data, len(rows), ln.ctypes.data, min_overlap))
    return rows


def sort_edges(edges: np.ndarray) -> np.ndarray:
    edges = np.ascontiguousarray(edges)
    _ck(lib().disco_host_sort_edges(edges.ctypes.data, len(edges)))
    return edges
