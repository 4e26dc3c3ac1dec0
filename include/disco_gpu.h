/*
 * disco_gpu.h -- C ABI of the B200 BuildGraph hot path (libdisco_gpu.so).
 *
 * The reference (abiswas-odu/Disco) has no plugin/FFI seam for this path: buildG is one process and the hot path is
 * the C++ call chain main.cpp:61-63
 *     HashTable::insertDataset()                      (src/BuildGraph/src/HashTable.cpp:46)
 *     OverlapGraph::buildOverlapGraphFromHashTable()  (src/BuildGraph/src/OverlapGraph.cpp:100)
 *        markContainedReads()                         (OverlapGraph.cpp:333)
 *        insertAllEdgesOfRead() / markTransitiveEdges() / removeTransitiveEdges()   (:631 / :687 / :731)
 * This header is the seam a maintainer would cut there (INTEGRATION.md shows the patch): the host keeps parsing,
 * filtering, numbering and file output; everything between "reads are numbered" and "edges are written" is one
 * library call.  Plain pointers and sizes only; no C++/torch types.
 *
 * Conventions
 *   - read ids are 0-based positions in the caller's accepted-read order (= reference readNumber - 1,
 *     Dataset.cpp:133-134 after file-order numbering).
 *   - packed reads use the reference record payload layout (HashTable.cpp:456-477): base i in bits
 *     [62-2*(i%32), 63-2*(i%32)] of word i/32, A=0 C=1 G=2 T=3, unused bits and words zero.
 *   - every function returns 0 on success or a negative DISCO_E_* code; disco_gpu_last_error() gives the text.
 *     Nothing in this library calls exit() or falls back to a CPU path: without a usable GPU every call fails.
 *   - a context is bound to one device and one stream; it is not thread-safe.
 */
#ifndef DISCO_GPU_H_
#define DISCO_GPU_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define DISCO_OK 0
#define DISCO_E_CUDA (-1)     /* CUDA runtime error (no device, launch failure, ...) */
#define DISCO_E_ARG (-2)      /* invalid argument / call order */
#define DISCO_E_NOMEM (-3)    /* device or host allocation failed */
#define DISCO_E_LIMIT (-4)    /* input exceeds a documented limit */

typedef struct disco_ctx disco_ctx;

/* One kept (transitively reduced) overlap edge, src < dst, written from src's point of view
 * == one line of <prefix>_<t>_parGraph.txt (OverlapGraph.cpp:808-833). */
typedef struct {
    uint32_t src;    /* 0-based read id */
    uint32_t dst;
    uint32_t offset; /* overlapOffset = len(src) - overlap length (OverlapGraph.cpp:667) */
    uint32_t orient; /* 0 u<-<v  1 u<->v  2 u>-<v  3 u>->v (Edge.h:30-34) */
} disco_edge;

/* One contained / duplicate read == one line of <prefix>_<t>_containedReads.txt (OverlapGraph.cpp:438-447). */
typedef struct {
    uint32_t contained; /* 0-based read id */
    uint32_t container;
    uint32_t orient;
    uint32_t start;     /* len(container) - overlapLen : where the contained read starts inside the container */
} disco_crow;

typedef struct {
    uint64_t n_reads, n_contained, n_edges;      /* n_edges = kept undirected edges */
    uint64_t raw_directed_edges;                  /* entries of the unreduced adjacency (both directions) */
    uint64_t cap_fired;                           /* (read, position) pairs where MAX_EDGE_PER_KMER cut candidates */
    uint64_t multi_overlap_pairs;                 /* pairs whose endpoints found different overlaps */
    uint64_t one_sided_edges;                     /* edges found from one endpoint only */
    uint64_t slow_path_reads;                     /* reads that took the sequential (exact cap) search path */
    uint64_t probes_contained, probes_edges;      /* k-mer look-ups issued */
    uint64_t buckets_contained, buckets_edges;    /* 32-byte table buckets read */
    uint64_t verified_contained, verified_edges;  /* candidate reads fetched and compared */
    uint64_t max_degree;
    uint64_t reduce_rows_fetched;                 /* neighbour rows read by the two reduction kernels */
    uint64_t reduce_entries_fetched;              /* adjacency entries read by the two reduction kernels */
    uint64_t table_buckets;                       /* size of the hash table in 32-byte buckets */
    uint64_t edge_capacity;                       /* adjacency entries allocated */
    uint64_t queries_contained, queries_edges;    /* reads searched by this context's launches */
    uint64_t kernel_launches;                     /* kernels this library launched since disco_gpu_begin (all contexts of the process) */
    uint64_t mark_rows_fetched, mark_entries_fetched; /* the marking kernel's share of reduce_rows/entries_fetched */
    /* device time of the last disco_gpu_build_graph(), CUDA events on the context's stream, milliseconds */
    float ms_table_all, ms_contained, ms_finish_contained, ms_table_nc, ms_edges, ms_mark, ms_emit, ms_total;
    float ms_edges_kernel, ms_contained_kernel; /* the search kernels alone (edge pass = probe + verify + exact) */
    float ms_edges_probe, ms_edges_verify, ms_edges_exact; /* the three kernels of the edge pass */
    float ms_mark_kernel, ms_emit_kernel;                   /* the two reduction kernels alone (ms_mark / ms_emit include
                                                             * whatever the caller did between the phases) */
} disco_stats;

/* One edge of the simplified graph (parsimplify's output, src/SimplifyGraph/src/OverlapGraphSimple.cpp:658-690): a maximal
 * chain of reduced edges through nodes with exactly one way in and one way out, written from its smaller end.  The
 * n_inner reads it swallowed are inner[inner_start ..], each (read | overlap offset << 32 | strand << 63) as in the
 * reference's packed list (EdgeSimple.cpp:226-230); edge length = offset_total + len(dst). */
typedef struct {
    uint32_t src, dst, orient, n_inner;
    uint64_t offset_total;
    uint64_t inner_start;
} disco_cedge;

/* number of CUDA devices this process can use (0 when there is none or the driver is missing); `buildG -g all` */
int disco_gpu_device_count(void);
/* ---- life cycle ---------------------------------------------------------------------------------------------- */
int disco_gpu_create(disco_ctx **out, int device);
void disco_gpu_destroy(disco_ctx *ctx);
const char *disco_gpu_last_error(const disco_ctx *ctx); /* ctx may be NULL: last error of a failed create */
/* Run on the caller's stream (a cudaStream_t passed as void*); default is a stream owned by the context. */
int disco_gpu_set_stream(disco_ctx *ctx, void *cuda_stream);

/* ---- input: replaces the record payload writes of HashTable::insertIntoTable (HashTable.cpp:456-514) --------- */
/* Host buffers (pinned memory makes the copy asynchronous). len[i] in (min_overlap, 32767]. */
int disco_gpu_load_reads(disco_ctx *ctx, const uint64_t *packed, const uint16_t *len, uint64_t n_reads,
                         uint32_t words_per_read);
/* Same, but the buffers already live in this device's memory. */
int disco_gpu_load_reads_device(disco_ctx *ctx, const uint64_t *d_packed, const uint16_t *d_len, uint64_t n_reads,
                                uint32_t words_per_read, uint32_t min_len, uint32_t max_len);
/* The caller's device buffers used in place, no copy, when the rows already have this library's pitch (the power of two
 * {2,4,8,16} words that holds the longest read, else an even word count; 16-byte aligned); any other pitch is copied as
 * above.  The buffers must stay valid and unchanged until the results have been read. */
int disco_gpu_use_reads_device(disco_ctx *ctx, const uint64_t *d_packed, const uint16_t *d_len, uint64_t n_reads,
                               uint32_t words_per_read, uint32_t min_len, uint32_t max_len);
/* disco_gpu_load_reads with the copy deferred into disco_gpu_build_graph (disco_gpu_phase_table(ctx, 0)): the rows cross
 * PCIe in chunks while the hash table is built from the chunks that have arrived (the reference fills its table while it
 * re-reads the files, HashTable.cpp:423-514).  The host buffers must stay valid and unchanged until that call returns.
 * min_len / max_len: shortest / longest read as the loader reports them (Dataset.cpp:139-146), or 0, 0 = find them here. */
int disco_gpu_load_reads_async(disco_ctx *ctx, const uint64_t *packed, const uint16_t *len, uint64_t n_reads,
                               uint32_t words_per_read, uint32_t min_len, uint32_t max_len);

/* ---- the whole hot path: insertDataset + buildOverlapGraphFromHashTable (HashTable.cpp:46, OverlapGraph.cpp:100)
 * min_overlap = MinOverlap4BuildGraph (main.cpp:170); max_edge_per_kmer = MAX_EDGE_PER_KMER (Common.h:62), 1..8. */
int disco_gpu_build_graph(disco_ctx *ctx, uint32_t min_overlap, uint32_t max_edge_per_kmer);

/* ---- results (device -> caller's host buffers) --------------------------------------------------------------- */
int disco_gpu_counts(disco_ctx *ctx, uint64_t *n_contained, uint64_t *n_edges);
/* rows in device order; the host library's sort helper (disco_host.h) restores the reference's -t 1 emission order */
int disco_gpu_get_contained(disco_ctx *ctx, disco_crow *rows, uint64_t capacity, uint64_t *n_written);
/* only the rows of the contained reads in [read_lo, read_hi): what one rank of a multi-GPU run writes (every context holds
 * all rows; BuildGraphMPI's ranks each write their own files, BuildGraphMPI/src/OverlapGraph.cpp:127, :370, :518) */
int disco_gpu_get_contained_range(disco_ctx *ctx, uint64_t read_lo, uint64_t read_hi, disco_crow *rows, uint64_t capacity, uint64_t *n_written);
/* edges in device emission order (callers sort if they need a canonical order) */
int disco_gpu_get_edges(disco_ctx *ctx, disco_edge *edges, uint64_t capacity, uint64_t *n_written);
/* Sorts the reduced edges by (src, dst) on the device: disco_gpu_get_edges then returns them in the order the parGraph
 * writer wants (replaces the host-side sort of ~10^7 records). */
int disco_gpu_sort_edges(disco_ctx *ctx);
/* Optional: a page-locked host buffer the emission kernel fills itself (over PCIe while it runs); disco_gpu_get_edges into
 * that same buffer then needs no copy.  NULL clears it; results larger than the capacity are not mirrored. */
int disco_gpu_set_edge_sink(disco_ctx *ctx, disco_edge *host_pinned, uint64_t capacity);
/* unreduced adjacency of one read (its own capped search, sorted by offset): for tests of the cap semantics */
int disco_gpu_get_row(disco_ctx *ctx, uint64_t read, disco_edge *out, uint64_t capacity, uint64_t *n_written);
int disco_gpu_get_stats(disco_ctx *ctx, disco_stats *out);

/* ---- the first consumer step, on the edges still in HBM: parsimplify's composite-edge contraction and
 * dead-end removal (OverlapGraphSimple.cpp:236-244: contract, then { contract; remove dead ends } until nothing changes).
 * min_overlap: edges with a shorter overlap are dropped first (:572); min_reads / min_len: a (composite) edge with at least
 * that many inner reads / that long keeps its end nodes from being dead ends (Config.cpp:43-44: 5 and 500).  Needs
 * disco_gpu_build_graph (or phase_reduce) to have run; single-GPU result sets (every edge in this context). */
int disco_gpu_simplify(disco_ctx *ctx, uint32_t min_overlap, uint32_t min_reads, uint32_t min_len,
                       uint64_t *n_edges, uint64_t *n_inner);
int disco_gpu_get_simplified(disco_ctx *ctx, disco_cedge *edges, uint64_t edge_capacity, uint64_t *inner, uint64_t inner_capacity);
/* rounds of contraction + dead-end removal, reduced edges removed with dead-end nodes, edges of closed chains (isolated
 * cycles: written unmerged -- the reference breaks them wherever its node order starts) and device milliseconds of the
 * last disco_gpu_simplify */
int disco_gpu_simplify_stats(disco_ctx *ctx, uint64_t *rounds, uint64_t *removed_edges, uint64_t *cycle_edges, float *ms);

/* ---- phase-level entry points (multi-GPU: query reads sharded by id, table and reads replicated; mirrors
 * BuildGraphMPI/src/OverlapGraph.cpp:524-529 and :293-295).  disco_gpu_build_graph() == the sequence
 *   table(0) -> contained(0,n) -> finish_contained -> table(1) -> edges(0,n) -> reduce(0,n).
 * table(1), the rebuild without the contained reads, is optional: when it is skipped the edge pass drops contained
 * candidates through the bitmap (same results, slower probes; worth it only when the table is replicated on many GPUs).
 * Between phases the host exchanges the buffers exposed below with NCCL. ---------------------------------------- */
int disco_gpu_begin(disco_ctx *ctx, uint32_t min_overlap, uint32_t max_edge_per_kmer);
int disco_gpu_phase_table(disco_ctx *ctx, int exclude_contained);
int disco_gpu_phase_contained(disco_ctx *ctx, uint64_t q_lo, uint64_t q_hi);
int disco_gpu_phase_finish_contained(disco_ctx *ctx);
int disco_gpu_phase_edges(disco_ctx *ctx, uint64_t q_lo, uint64_t q_hi);
/* the same pass in pieces: parts of [q_lo,q_hi) in ascending order, the first one starting at q_lo; rows append */
int disco_gpu_phase_edges_part(disco_ctx *ctx, uint64_t q_lo, uint64_t q_hi, uint64_t part_lo, uint64_t part_hi);
int disco_gpu_phase_reduce(disco_ctx *ctx, uint64_t u_lo, uint64_t u_hi);
/* the two halves of phase_reduce: markTransitiveEdges (OverlapGraph.cpp:687-723) and removeTransitiveEdges + emission
 * (:731-761, :808).  Separate so that a multi-GPU caller can put a barrier between them (emit reads the marks of
 * neighbouring nodes, which another GPU may own). */
int disco_gpu_phase_reduce_mark(disco_ctx *ctx, uint64_t u_lo, uint64_t u_hi);
int disco_gpu_phase_reduce_emit(disco_ctx *ctx, uint64_t u_lo, uint64_t u_hi);
/* device pointers for collectives: containment keys u64[n] (all-reduce MIN), row info u64[n] (all-reduce SUM after
 * rebase), adjacency entries u64[*n_entries] (all-gather) */
void *disco_gpu_dev_contained_keys(disco_ctx *ctx);
void *disco_gpu_dev_rowinfo(disco_ctx *ctx);
void *disco_gpu_dev_rows(disco_ctx *ctx, uint64_t *n_entries);
/* Sparse exchange of the containment keys (only a few percent of the reads are contained): compact_keys writes this
 * rank's keys that are set as u64 pairs (read, key) into the caller's device buffer (capacity in pairs) and returns their
 * number (larger than the capacity: nothing usable was written -- all-reduce the dense array instead); apply_keys takes
 * the minimum over pairs gathered from every rank (a pair whose read is >= n is padding). */
int disco_gpu_compact_keys(disco_ctx *ctx, void *d_pairs, uint64_t capacity, uint64_t *n_pairs);
int disco_gpu_apply_keys(disco_ctx *ctx, const void *d_pairs, uint64_t n_pairs);
/* Adjacency exchange helpers.  Common layout on every rank: rank r's rows live at [r * slot, r * slot + count_r).
 *   reserve_rows : make the adjacency buffer hold at least n_entries (contents kept)
 *   move_rows    : move this rank's rows from the front of the buffer to dst_offset (regions may not overlap)
 *   rebase_rows  : add `base` to the start of every local row in [u_lo,u_hi) (before the row-info all-reduce)
 *   set_rows_used: entries in use after the gather / where the next part of the edge pass starts appending
 *   use_rows     : let the reduction read a caller-owned device buffer (the gathered adjacency) instead, no copy
 *   adopt_rows   : alternative: replace the adjacency by a copy of d_rows[n_entries] */
int disco_gpu_reserve_rows(disco_ctx *ctx, uint64_t n_entries);
int disco_gpu_move_rows(disco_ctx *ctx, uint64_t dst_offset);
int disco_gpu_rebase_rows(disco_ctx *ctx, uint64_t u_lo, uint64_t u_hi, uint64_t base);
int disco_gpu_set_rows_used(disco_ctx *ctx, uint64_t n_entries);
int disco_gpu_use_rows(disco_ctx *ctx, const uint64_t *d_rows, uint64_t n_entries);
int disco_gpu_adopt_rows(disco_ctx *ctx, const uint64_t *d_rows, uint64_t n_entries);
/* ---- key-sharded table and range-partitioned adjacency (Mode B; the partitioning of BuildGraphMPIRMA,
 * src/BuildGraphMPIRMA/src/HashTable.cpp: each rank owns a slice of the hash table, the others reach it by one-sided
 * MPI_Get).  Reads stay replicated.  Shard = mulhi(fingerprint, world): every GPU scans all reads, inserts the keys of
 * its own shard (no communication) and sets every key's bit in its own copy of the presence filter.  Probes of another
 * shard are plain loads through NVLink peer pointers (CUDA IPC mappings) issued by the same kernels -- the one-sided
 * get of the reference, without a host in the loop.  The adjacency is not gathered either: a rank keeps the rows of its
 * own query range and the reduction reads neighbours' rows from their owner the same way.
 *   set_shard    : before disco_gpu_begin; world = 1 returns to the single-table mode
 *   export_mem   : CUDA IPC handle (DISCO_IPC_HANDLE_BYTES) of this rank's table shard / adjacency buffer
 *   import_peers : handles of all ranks, [world] x DISCO_IPC_HANDLE_BYTES in rank order (own entry ignored); for the
 *                  adjacency also the read-id bounds u64[world + 1] of the ranks' query ranges.  Call again whenever a
 *                  rank's buffer may have been reallocated (mappings of unchanged handles are kept).
 * The caller synchronises the ranks (barrier) between the phases: table complete before anybody probes, probing
 * finished before the table is rebuilt, edge pass finished before the reduction, marks finished before the emission. */
#define DISCO_MAX_SHARDS 8
#define DISCO_IPC_HANDLE_BYTES 64
#define DISCO_MEM_TABLE 0
#define DISCO_MEM_ROWS 1
int disco_gpu_set_shard(disco_ctx *ctx, uint32_t world, uint32_t rank);
/* The same with the choice of what is partitioned: shard_table != 0 is disco_gpu_set_shard; shard_table == 0 keeps the
 * hash table replicated (every GPU builds it from all reads and probes it locally, as BuildGraphMPI does) and partitions
 * only the adjacency by query range -- neighbours' rows are then read from their owner through peer pointers instead of
 * being all-gathered (import DISCO_MEM_ROWS only). */
int disco_gpu_set_partition(disco_ctx *ctx, uint32_t world, uint32_t rank, int shard_table);
int disco_gpu_export_mem(disco_ctx *ctx, int which, void *handle_out);
int disco_gpu_import_peers(disco_ctx *ctx, int which, const void *handles, const uint64_t *bounds);
/* shards held by contexts of the same process: device pointers [world] instead of IPC handles (the table shard from
 * disco_gpu_dev_table, the adjacency from disco_gpu_dev_rows); peer access between the devices is the caller's job */
int disco_gpu_import_peer_ptrs(disco_ctx *ctx, int which, const void *const *device_ptrs, const uint64_t *bounds);
void *disco_gpu_dev_table(disco_ctx *ctx);
/* Caller-owned device memory instead of the library's own: the table shard (after disco_gpu_begin; at least
 * disco_gpu_table_words() u64) or the adjacency (before the edge pass; capacity in entries -- never grown by the
 * library, the edge pass returns DISCO_E_NOMEM when it is too small).  Meant for memory the peers map by other means
 * than CUDA IPC handles, e.g. symmetric memory (VMM allocations, 2 MB pages: legacy IPC mappings thrash the requester's
 * TLB once the remote footprint exceeds ~1-2 GB). */
uint64_t disco_gpu_table_words(disco_ctx *ctx);
int disco_gpu_adopt_buffer(disco_ctx *ctx, int which, void *d_ptr, uint64_t n_u64);
/* The whole of Mode B for several GPUs driven by ONE process (`buildG -g 0,1,...`): one context per GPU (or several on
 * one GPU), every context holding the same reads; one host thread per context, peer access instead of IPC, the
 * containment keys min-reduced and the row infos copied over NVLink by the library itself -- no NCCL, no MPI.
 * Afterwards context r holds the edges whose lower endpoint lies in [r*n/world, (r+1)*n/world) and every context holds
 * all contained rows.  Returns the first failing rank's code (its message: disco_gpu_last_error of that context). */
int disco_gpu_build_graph_multi(disco_ctx *const *ctxs, uint32_t world, uint32_t min_overlap, uint32_t max_edge_per_kmer);
/* largest row length over all ranks (sizes the reduction kernel's shared memory) */
int disco_gpu_set_max_degree(disco_ctx *ctx, uint64_t max_degree);
/* wait for everything queued on the context's stream */
int disco_gpu_sync(disco_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* DISCO_GPU_H_ */
