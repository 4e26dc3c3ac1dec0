"""ctypes binding of the C ABI in include/disco_gpu.h (libdisco_gpu.so).

This is the only way Python reaches the kernels; there is no fallback: if the library is missing or no GPU is
present the calls raise."""
import ctypes as C
import os
import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libdisco_gpu.so")

EDGE_DTYPE = np.dtype([("src", "<u4"), ("dst", "<u4"), ("offset", "<u4"), ("orient", "<u4")])
CROW_DTYPE = np.dtype([("contained", "<u4"), ("container", "<u4"), ("orient", "<u4"), ("start", "<u4")])
CEDGE_DTYPE = np.dtype([("src", "<u4"), ("dst", "<u4"), ("orient", "<u4"), ("n_inner", "<u4"), ("offset_total", "<u8"), ("inner_start", "<u8")])

EXPORTS = [
    "disco_gpu_create", "disco_gpu_destroy", "disco_gpu_last_error", "disco_gpu_set_stream", "disco_gpu_load_reads",
    "disco_gpu_load_reads_device", "disco_gpu_build_graph", "disco_gpu_counts", "disco_gpu_get_contained",
    "disco_gpu_get_edges", "disco_gpu_get_row", "disco_gpu_get_stats", "disco_gpu_begin", "disco_gpu_phase_table",
    "disco_gpu_phase_contained", "disco_gpu_phase_finish_contained", "disco_gpu_phase_edges", "disco_gpu_phase_reduce",
    "disco_gpu_dev_contained_keys", "disco_gpu_dev_rowinfo", "disco_gpu_dev_rows", "disco_gpu_rebase_rows",
    "disco_gpu_adopt_rows", "disco_gpu_set_max_degree", "disco_gpu_sync", "disco_gpu_reserve_rows", "disco_gpu_move_rows",
    "disco_gpu_set_rows_used", "disco_gpu_use_rows", "disco_gpu_phase_edges_part", "disco_gpu_phase_reduce_mark",
    "disco_gpu_phase_reduce_emit", "disco_gpu_set_shard", "disco_gpu_export_mem", "disco_gpu_import_peers",
    "disco_gpu_import_peer_ptrs", "disco_gpu_dev_table", "disco_gpu_table_words", "disco_gpu_adopt_buffer",
    "disco_gpu_build_graph_multi", "disco_gpu_device_count", "disco_gpu_set_partition", "disco_gpu_compact_keys", "disco_gpu_apply_keys", "disco_gpu_simplify", "disco_gpu_get_simplified", "disco_gpu_simplify_stats", "disco_gpu_set_edge_sink",
    "disco_gpu_use_reads_device", "disco_gpu_load_reads_async", "disco_gpu_sort_edges", "disco_gpu_get_contained_range",
]
MAX_SHARDS, IPC_HANDLE_BYTES, MEM_TABLE, MEM_ROWS = 8, 64, 0, 1


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "n_reads", "n_contained", "n_edges", "raw_directed_edges", "cap_fired", "multi_overlap_pairs", "one_sided_edges",
        "slow_path_reads", "probes_contained", "probes_edges", "buckets_contained", "buckets_edges",
        "verified_contained", "verified_edges", "max_degree", "reduce_rows_fetched", "reduce_entries_fetched",
        "table_buckets", "edge_capacity", "queries_contained", "queries_edges", "kernel_launches", "mark_rows_fetched", "mark_entries_fetched")] + [(n, C.c_float) for n in (
            "ms_table_all", "ms_contained", "ms_finish_contained", "ms_table_nc", "ms_edges", "ms_mark", "ms_emit", "ms_total",
            "ms_edges_kernel", "ms_contained_kernel", "ms_edges_probe", "ms_edges_verify", "ms_edges_exact",
            "ms_mark_kernel", "ms_emit_kernel")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class DiscoError(RuntimeError):
    pass


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DiscoError(f"{LIB_PATH} not built: run `python -m disco_b200.build` (there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        vp, u64, u32, i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int
        L.disco_gpu_create.argtypes = [C.POINTER(vp), i32]
        L.disco_gpu_destroy.argtypes = [vp]
        L.disco_gpu_destroy.restype = None
        L.disco_gpu_last_error.argtypes = [vp]
        L.disco_gpu_last_error.restype = C.c_char_p
        L.disco_gpu_set_stream.argtypes = [vp, vp]
        L.disco_gpu_load_reads.argtypes = [vp, vp, vp, u64, u32]
        L.disco_gpu_load_reads_device.argtypes = [vp, vp, vp, u64, u32, u32, u32]
        L.disco_gpu_use_reads_device.argtypes = [vp, vp, vp, u64, u32, u32, u32]
        L.disco_gpu_load_reads_async.argtypes = [vp, vp, vp, u64, u32, u32, u32]
        L.disco_gpu_build_graph.argtypes = [vp, u32, u32]
        L.disco_gpu_counts.argtypes = [vp, C.POINTER(u64), C.POINTER(u64)]
        L.disco_gpu_get_contained.argtypes = [vp, vp, u64, C.POINTER(u64)]
        L.disco_gpu_get_contained_range.argtypes = [vp, u64, u64, vp, u64, C.POINTER(u64)]
        L.disco_gpu_get_edges.argtypes = [vp, vp, u64, C.POINTER(u64)]
        L.disco_gpu_get_row.argtypes = [vp, u64, vp, u64, C.POINTER(u64)]
        L.disco_gpu_get_stats.argtypes = [vp, C.POINTER(Stats)]
        L.disco_gpu_begin.argtypes = [vp, u32, u32]
        L.disco_gpu_phase_table.argtypes = [vp, i32]
        L.disco_gpu_phase_contained.argtypes = [vp, u64, u64]
        L.disco_gpu_phase_finish_contained.argtypes = [vp]
        L.disco_gpu_phase_edges.argtypes = [vp, u64, u64]
        L.disco_gpu_phase_reduce.argtypes = [vp, u64, u64]
        L.disco_gpu_phase_reduce_mark.argtypes = [vp, u64, u64]
        L.disco_gpu_phase_reduce_emit.argtypes = [vp, u64, u64]
        L.disco_gpu_set_shard.argtypes = [vp, u32, u32]
        L.disco_gpu_set_partition.argtypes = [vp, u32, u32, i32]
        L.disco_gpu_compact_keys.argtypes = [vp, vp, u64, C.POINTER(u64)]
        L.disco_gpu_set_edge_sink.argtypes = [vp, vp, u64]
        L.disco_gpu_sort_edges.argtypes = [vp]
        L.disco_gpu_simplify.argtypes = [vp, u32, u32, u32, C.POINTER(u64), C.POINTER(u64)]
        L.disco_gpu_get_simplified.argtypes = [vp, vp, u64, vp, u64]
        L.disco_gpu_simplify_stats.argtypes = [vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64), C.POINTER(C.c_float)]
        L.disco_gpu_apply_keys.argtypes = [vp, vp, u64]
        L.disco_gpu_export_mem.argtypes = [vp, i32, vp]
        L.disco_gpu_import_peers.argtypes = [vp, i32, vp, vp]
        L.disco_gpu_import_peer_ptrs.argtypes = [vp, i32, vp, vp]
        L.disco_gpu_dev_table.argtypes = [vp]
        L.disco_gpu_dev_table.restype = vp
        L.disco_gpu_table_words.argtypes = [vp]
        L.disco_gpu_table_words.restype = u64
        L.disco_gpu_adopt_buffer.argtypes = [vp, i32, vp, u64]
        L.disco_gpu_build_graph_multi.argtypes = [vp, u32, u32, u32]
        L.disco_gpu_device_count.argtypes = []
        for f in ("disco_gpu_dev_contained_keys", "disco_gpu_dev_rowinfo"):
            getattr(L, f).argtypes = [vp]
            getattr(L, f).restype = vp
        L.disco_gpu_dev_rows.argtypes = [vp, C.POINTER(u64)]
        L.disco_gpu_dev_rows.restype = vp
        L.disco_gpu_rebase_rows.argtypes = [vp, u64, u64, u64]
        L.disco_gpu_adopt_rows.argtypes = [vp, vp, u64]
        L.disco_gpu_set_max_degree.argtypes = [vp, u64]
        L.disco_gpu_reserve_rows.argtypes = [vp, u64]
        L.disco_gpu_move_rows.argtypes = [vp, u64]
        L.disco_gpu_set_rows_used.argtypes = [vp, u64]
        L.disco_gpu_use_rows.argtypes = [vp, vp, u64]
        L.disco_gpu_phase_edges_part.argtypes = [vp, u64, u64, u64, u64]
        L.disco_gpu_sync.argtypes = [vp]
        _lib = L
    return _lib


class GpuBuildGraph:
    """One context = one GPU.  Mirrors the reference call order: load reads (HashTable::insertDataset) then
    build_graph (OverlapGraph::buildOverlapGraphFromHashTable)."""

    def __init__(self, device: int = 0):
        self._L = lib()
        self._h = C.c_void_p()
        rc = self._L.disco_gpu_create(C.byref(self._h), device)
        if rc:
            raise DiscoError(f"disco_gpu_create({device}) -> {rc}: {self._L.disco_gpu_last_error(None).decode()}")
        self.n = 0
        self._keep = None

    def close(self):
        if self._h:
            self._L.disco_gpu_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc:
            raise DiscoError(f"{what} -> {rc}: {self._L.disco_gpu_last_error(self._h).decode()}")

    def set_stream(self, cuda_stream_handle: int):
        self._ck(self._L.disco_gpu_set_stream(self._h, C.c_void_p(cuda_stream_handle)), "set_stream")

    def sync(self):
        self._ck(self._L.disco_gpu_sync(self._h), "sync")

    def load_reads(self, packed: np.ndarray, lens: np.ndarray):
        """packed uint64[n, wpr] (host, ideally pinned), lens uint16[n]."""
        assert packed.dtype == np.uint64 and packed.ndim == 2 and packed.flags.c_contiguous
        assert lens.dtype == np.uint16 and lens.shape == (packed.shape[0],)
        self._keep = (packed, lens)  # the copy is asynchronous
        self.n = packed.shape[0]
        self._ck(self._L.disco_gpu_load_reads(self._h, packed.ctypes.data, lens.ctypes.data, self.n, packed.shape[1]), "load_reads")

    def load_reads_ptr(self, packed_ptr: int, lens_ptr: int, n: int, wpr: int):
        self.n = n
        self._ck(self._L.disco_gpu_load_reads(self._h, C.c_void_p(packed_ptr), C.c_void_p(lens_ptr), n, wpr), "load_reads")

    def load_reads_device(self, d_packed_ptr: int, d_lens_ptr: int, n: int, wpr: int, min_len: int, max_len: int):
        self.n = n
        self._ck(self._L.disco_gpu_load_reads_device(self._h, C.c_void_p(d_packed_ptr), C.c_void_p(d_lens_ptr), n, wpr,
                                                     min_len, max_len), "load_reads_device")

    def use_reads_device(self, d_packed_ptr: int, d_lens_ptr: int, n: int, wpr: int, min_len: int, max_len: int):
        """the caller's device buffers in place (no copy) when wpr is the library's row pitch; they must outlive the run"""
        self.n = n
        self._ck(self._L.disco_gpu_use_reads_device(self._h, C.c_void_p(d_packed_ptr), C.c_void_p(d_lens_ptr), n, wpr,
                                                    min_len, max_len), "use_reads_device")

    def load_reads_async(self, packed_ptr: int, lens_ptr: int, n: int, wpr: int, min_len: int = 0, max_len: int = 0):
        """host buffers (pinned) whose upload is deferred into build_graph and overlapped with the table build; they must
        stay valid and unchanged until build_graph has returned"""
        self.n = n
        self._ck(self._L.disco_gpu_load_reads_async(self._h, C.c_void_p(packed_ptr), C.c_void_p(lens_ptr), n, wpr, min_len, max_len),
                 "load_reads_async")

    def build_graph(self, min_overlap: int, max_edge_per_kmer: int = 4):
        self._ck(self._L.disco_gpu_build_graph(self._h, min_overlap, max_edge_per_kmer), "build_graph")

    # phase level
    def begin(self, min_overlap, max_edge_per_kmer=4):
        self._ck(self._L.disco_gpu_begin(self._h, min_overlap, max_edge_per_kmer), "begin")

    def phase_table(self, exclude_contained: bool):
        self._ck(self._L.disco_gpu_phase_table(self._h, int(exclude_contained)), "phase_table")

    def phase_contained(self, lo, hi):
        self._ck(self._L.disco_gpu_phase_contained(self._h, lo, hi), "phase_contained")

    def phase_finish_contained(self):
        self._ck(self._L.disco_gpu_phase_finish_contained(self._h), "phase_finish_contained")

    def phase_edges(self, lo, hi):
        self._ck(self._L.disco_gpu_phase_edges(self._h, lo, hi), "phase_edges")

    def phase_edges_part(self, lo, hi, part_lo, part_hi):
        self._ck(self._L.disco_gpu_phase_edges_part(self._h, lo, hi, part_lo, part_hi), "phase_edges_part")

    def use_rows(self, d_rows_ptr: int, n_entries: int):
        self._ck(self._L.disco_gpu_use_rows(self._h, C.c_void_p(d_rows_ptr), n_entries), "use_rows")

    def phase_reduce(self, lo, hi):
        self._ck(self._L.disco_gpu_phase_reduce(self._h, lo, hi), "phase_reduce")

    def phase_reduce_mark(self, lo, hi):
        self._ck(self._L.disco_gpu_phase_reduce_mark(self._h, lo, hi), "phase_reduce_mark")

    def phase_reduce_emit(self, lo, hi):
        self._ck(self._L.disco_gpu_phase_reduce_emit(self._h, lo, hi), "phase_reduce_emit")

    # key-sharded mode (Mode B)
    def set_shard(self, world: int, rank: int):
        self._ck(self._L.disco_gpu_set_shard(self._h, world, rank), "set_shard")

    def set_partition(self, world: int, rank: int, shard_table: bool):
        self._ck(self._L.disco_gpu_set_partition(self._h, world, rank, int(shard_table)), "set_partition")

    def export_mem(self, which: int) -> bytes:
        buf = C.create_string_buffer(IPC_HANDLE_BYTES)
        self._ck(self._L.disco_gpu_export_mem(self._h, which, buf), "export_mem")
        return buf.raw

    def import_peers(self, which: int, handles, bounds=None):
        """handles: one IPC_HANDLE_BYTES blob per rank, in rank order; bounds: world + 1 read ids (adjacency only)"""
        blob = b"".join(bytes(h) for h in handles)
        b = (C.c_uint64 * len(bounds))(*bounds) if bounds is not None else None
        self._ck(self._L.disco_gpu_import_peers(self._h, which, blob, b), "import_peers")

    def import_peer_ptrs(self, which: int, ptrs, bounds=None):
        """the same for shards owned by contexts of this process: device pointers in rank order"""
        arr = (C.c_void_p * len(ptrs))(*ptrs)
        b = (C.c_uint64 * len(bounds))(*bounds) if bounds is not None else None
        self._ck(self._L.disco_gpu_import_peer_ptrs(self._h, which, arr, b), "import_peer_ptrs")

    def dev_table(self) -> int:
        return self._L.disco_gpu_dev_table(self._h)

    def table_words(self) -> int:
        return self._L.disco_gpu_table_words(self._h)

    def adopt_buffer(self, which: int, d_ptr: int, n_u64: int):
        self._ck(self._L.disco_gpu_adopt_buffer(self._h, which, C.c_void_p(d_ptr), n_u64), "adopt_buffer")

    def compact_keys(self, d_pairs_ptr: int, capacity: int) -> int:
        n = C.c_uint64()
        self._ck(self._L.disco_gpu_compact_keys(self._h, C.c_void_p(d_pairs_ptr), capacity, C.byref(n)), "compact_keys")
        return n.value

    def apply_keys(self, d_pairs_ptr: int, n_pairs: int):
        self._ck(self._L.disco_gpu_apply_keys(self._h, C.c_void_p(d_pairs_ptr), n_pairs), "apply_keys")

    def dev_contained_keys(self) -> int:
        return self._L.disco_gpu_dev_contained_keys(self._h)

    def dev_rowinfo(self) -> int:
        return self._L.disco_gpu_dev_rowinfo(self._h)

    def dev_rows(self):
        n = C.c_uint64()
        p = self._L.disco_gpu_dev_rows(self._h, C.byref(n))
        return p, n.value

    def rebase_rows(self, lo, hi, base):
        self._ck(self._L.disco_gpu_rebase_rows(self._h, lo, hi, base), "rebase_rows")

    def adopt_rows(self, d_rows_ptr: int, n_entries: int):
        self._ck(self._L.disco_gpu_adopt_rows(self._h, C.c_void_p(d_rows_ptr), n_entries), "adopt_rows")

    def reserve_rows(self, n_entries: int):
        self._ck(self._L.disco_gpu_reserve_rows(self._h, n_entries), "reserve_rows")

    def move_rows(self, dst_offset: int):
        self._ck(self._L.disco_gpu_move_rows(self._h, dst_offset), "move_rows")

    def set_rows_used(self, n_entries: int):
        self._ck(self._L.disco_gpu_set_rows_used(self._h, n_entries), "set_rows_used")

    def set_max_degree(self, d: int):
        self._ck(self._L.disco_gpu_set_max_degree(self._h, d), "set_max_degree")

    # results
    def counts(self):
        a, b = C.c_uint64(), C.c_uint64()
        self._ck(self._L.disco_gpu_counts(self._h, C.byref(a), C.byref(b)), "counts")
        return a.value, b.value

    def contained(self) -> np.ndarray:
        nc, _ = self.counts()
        out = np.zeros(nc, dtype=CROW_DTYPE)
        w = C.c_uint64()
        self._ck(self._L.disco_gpu_get_contained(self._h, out.ctypes.data, nc, C.byref(w)), "get_contained")
        return out[:w.value]

    def contained_into(self, out: np.ndarray) -> np.ndarray:
        w = C.c_uint64()
        self._ck(self._L.disco_gpu_get_contained(self._h, out.ctypes.data, len(out), C.byref(w)), "get_contained")
        return out[:w.value]

    def contained_range_into(self, out: np.ndarray, read_lo: int, read_hi: int) -> np.ndarray:
        """the rows of the contained reads in [read_lo, read_hi) only (a rank's share of a multi-GPU run)"""
        w = C.c_uint64()
        self._ck(self._L.disco_gpu_get_contained_range(self._h, read_lo, read_hi, out.ctypes.data, len(out), C.byref(w)), "get_contained_range")
        return out[:w.value]

    def edges(self, out: np.ndarray = None) -> np.ndarray:
        _, ne = self.counts()
        if out is None:
            out = np.zeros(ne, dtype=EDGE_DTYPE)
        w = C.c_uint64()
        self._ck(self._L.disco_gpu_get_edges(self._h, out.ctypes.data, len(out), C.byref(w)), "get_edges")
        return out[:w.value]

    def sort_edges(self):
        """sort the reduced edges by (src, dst) on the device; edges() then returns them in file order"""
        self._ck(self._L.disco_gpu_sort_edges(self._h), "sort_edges")

    def set_edge_sink(self, host_ptr: int, capacity: int):
        """pinned host buffer (EDGE_DTYPE[capacity]) the emission kernel fills while it runs; 0 clears it"""
        self._ck(self._L.disco_gpu_set_edge_sink(self._h, C.c_void_p(host_ptr), capacity), "set_edge_sink")

    def simplify(self, min_overlap: int = 0, min_reads: int = 5, min_len: int = 500):
        """parsimplify's composite-edge contraction + dead-end removal on the device-resident reduced edges (defaults:
        Config.cpp:43-44).  Returns (edges CEDGE_DTYPE, inner u64 [read | offset << 32 | strand << 63], stats dict)."""
        ne, ni = C.c_uint64(), C.c_uint64()
        self._ck(self._L.disco_gpu_simplify(self._h, min_overlap, min_reads, min_len, C.byref(ne), C.byref(ni)), "simplify")
        e = np.zeros(ne.value, dtype=CEDGE_DTYPE)
        inner = np.zeros(ni.value, dtype=np.uint64)
        self._ck(self._L.disco_gpu_get_simplified(self._h, e.ctypes.data, len(e), inner.ctypes.data, len(inner)), "get_simplified")
        r, rm, cy, ms = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_float()
        self._ck(self._L.disco_gpu_simplify_stats(self._h, C.byref(r), C.byref(rm), C.byref(cy), C.byref(ms)), "simplify_stats")
        return e, inner, {"rounds": r.value, "removed_edges": rm.value, "cycle_edges": cy.value, "ms": ms.value}

    def row(self, read: int, capacity: int = 1 << 16) -> np.ndarray:
        out = np.zeros(capacity, dtype=EDGE_DTYPE)
        w = C.c_uint64()
        self._ck(self._L.disco_gpu_get_row(self._h, read, out.ctypes.data, capacity, C.byref(w)), "get_row")
        return out[:w.value]

    def stats(self) -> dict:
        s = Stats()
        self._ck(self._L.disco_gpu_get_stats(self._h, C.byref(s)), "get_stats")
        return s.as_dict()


def device_count() -> int:
    return lib().disco_gpu_device_count()


def build_graph_multi(graphs, min_overlap: int, max_edge_per_kmer: int = 4):
    """Mode B over several contexts of THIS process (one per GPU, or several on one GPU), all holding the same reads:
    one host thread per context inside the library, peer access between the devices, no NCCL.  Afterwards graphs[r]
    holds the edges whose lower endpoint lies in rank r's read range; every context holds all contained rows."""
    L = lib()
    arr = (C.c_void_p * len(graphs))(*[g._h for g in graphs])
    rc = L.disco_gpu_build_graph_multi(arr, len(graphs), min_overlap, max_edge_per_kmer)
    if rc:
        msgs = [L.disco_gpu_last_error(g._h).decode() for g in graphs]
        raise DiscoError(f"build_graph_multi -> {rc}: " + " | ".join(m for m in msgs if m))

def sort_edges(e: np.ndarray) -> np.ndarray:
    return e[np.lexsort((e["orient"], e["offset"], e["dst"], e["src"]))]
