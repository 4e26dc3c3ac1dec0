// kernels.cuh -- parameter blocks and launcher prototypes of the sm_100a kernels (kernels.cu) used by api.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/disco_gpu.h"

namespace disco {

// per-launch statistics slots (unsigned long long each)
enum StatSlot {
    ST_QUERIES = 0,   // reads searched
    ST_PROBES,        // k-mer look-ups
    ST_BUCKETS,       // 32-byte buckets read
    ST_VERIFIED,      // candidate reads fetched + compared
    ST_HITS,          // verified candidates that passed
    ST_ENTRIES,       // adjacency entries written (directed)
    ST_CAP_FIRED,
    ST_SLOW_READS,
    ST_MAXDEG,        // atomicMax
    ST_OVERFLOW,      // bit 0: adjacency buffer too small (entries needed = rows cursor); bit 1: candidate buffer (its cursor)
    ST_ROWS_FETCHED,  // reduction: neighbour rows read
    ST_ENTRIES_FETCHED,
    ST_MULTI_OVERLAP, // from here on: written by the emission kernel only (zeroed before each of its attempts)
    ST_ONE_SIDED,
    ST_EDGES_OUT,
    ST_EMIT_ROWS,
    ST_EMIT_ENTRIES,
    ST_COUNT
};

struct ReadsView {
    const uint64_t *words; // n * stride, 16-byte aligned rows
    const uint64_t *words_rc; // the reverse complement of every read in the same layout, or null (see k_edges_verify)
    const uint64_t *tails;    // u64[n][4]: the last 128 bases of every read (zero padded in front), or null (see k_verify_flat)
    const uint16_t *len;   // n
    uint64_t n;
    int stride;            // words per read (even)
    int uniform_len;       // 0, or the common length of every read
    int min_len, max_len;
};

struct TableView {
    uint64_t *slots;   // nbuckets * 4 (this GPU's shard when world > 1)
    uint64_t nbuckets; // buckets per shard
    uint32_t *filter;  // presence bitmap over a second slice of the k-mer hash, small enough to stay in L2 (or null)
    uint32_t filter_mask; // bits - 1 (power of two)
    unsigned int *full;   // set by the insert kernels when a walk went round the whole (shard of the) table
    // key-sharded table (Mode B, BuildGraphMPIRMA's partitioning): shard = mulhi(hash, world); every GPU inserts the
    // keys of its own shard and probes the other shards through NVLink peer pointers.  world <= 1: single table.
    const uint64_t *const *peers; // device array [world] of shard base pointers (peer-mapped; [rank] == slots)
    uint32_t world, rank;
};

// Binned table build: the records (fingerprint, slot value) sorted by bucket range into `nbins` bins before they are
// inserted, so that the slice of the table a bin maps to is filled while it is L2 resident (k_table_bin / k_table_fill)
struct BinView {
    ulonglong2 *recs;             // nbins * cap records
    unsigned long long *count;    // records per bin (keeps counting past cap: then *overflow is set)
    uint64_t cap;                 // records a bin holds
    uint32_t nbins;
    unsigned int *overflow;       // some bin was too small (skewed k-mers): the direct insert kernel takes over
};

struct SearchParams {
    ReadsView reads;
    TableView table;
    int K;                 // hashStringLength = minOverlap - 1
    int cap;               // MAX_EDGE_PER_KMER
    uint64_t q_lo, q_hi;   // query reads [q_lo, q_hi)
    unsigned long long *work_counter; // [0] (and [1], [2] for the verify / exact kernels of the edge pass)
    unsigned long long *stats;
    // containment pass
    unsigned long long *best; // n keys, ~0 = not contained
    // edge pass
    const uint32_t *contained_bits;
    int only_flagged;   // containment, warp-per-read kernel as the fall-back of the flat one: only reads whose row info says so
    int skip_contained; // the table still holds contained reads (built once): drop them as candidates (HashTable.cpp:533)
    uint64_t *rows;
    uint64_t rows_cap;
    unsigned long long *rows_cursor;
    uint64_t *rowinfo;
    // flat edge pass (edges_flat.cuh): candidate list of every 32-read batch, in up to 4 segments per batch
    uint64_t *cands;
    uint64_t cands_cap;
    unsigned long long *cands_cursor;
    uint64_t *batchinfo; // 4 u64 per batch of the launch: (first candidate << 20) | count
    // shared-memory shape, filled in by the launcher
    int dbg;     // timing ablations (env DISCO_DBG, results are then wrong): 1 no compare, 2 stop after hashing, 4 stop after probing
    int npos;    // k-mer positions of the longest read (max_len - K)
    int hcap;    // candidate queue entries per warp
    int hset;    // neighbour-id set slots (u32, power of two)
    int rowcap;  // row buffer entries per warp (>= cap * (max_len - K), also holds the fast path's scratch)
};

struct ReduceParams {
    ReadsView reads;
    uint64_t *rows;
    const uint64_t *rowinfo;
    // range-partitioned adjacency (Mode B): the row of read v lives on the GPU whose [bounds[r], bounds[r+1]) holds v,
    // at peer_rows[r] + start(v).  world <= 1: every row is in `rows`.
    const uint64_t *const *peer_rows; // device array [world]
    const uint64_t *bounds;           // device array [world + 1]
    uint32_t world;
    int peer_load; // how remote rows are read (env DISCO_PEER_LOAD, tuning; no measurable difference): 0 ld.cg, 1 ld.nc
    uint64_t u_lo, u_hi;
    unsigned long long *work_counter;
    unsigned long long *stats;
    int maxdeg;
    // emit
    void *edges_out;       // disco_edge[edges_cap]
    uint64_t edges_cap;
    void *edges_out2;      // optional second destination (the caller's pinned host buffer, written over PCIe while the kernel runs)
    uint64_t edges_cap2;
    unsigned long long *edges_cursor;
};

// (r_lo, r_hi: the reads to insert -- the chunked upload of api.cu inserts each chunk as it arrives; default: all.
//  gate: the kernel only runs when *gate != 0 -- the fallback behind a binned build whose bins overflowed)
cudaError_t launch_table_insert(const ReadsView &r, const TableView &t, int K, const uint32_t *skip_bits,
                                int num_sms, cudaStream_t s, uint64_t r_lo = 0, uint64_t r_hi = ~0ULL, const unsigned int *gate = nullptr);
// binned build: hash + bin the reads [r_lo, r_hi) (sets the filter bits), then fill the table bin by bin (skip_bits: leave
// out the records of contained reads; set_filter: the filter was cleared and is rebuilt from the records that are kept)
bool table_bin_supported(int max_len);
cudaError_t launch_table_bin(const ReadsView &r, const TableView &t, int K, const BinView &b, int num_sms, cudaStream_t s,
                             uint64_t r_lo = 0, uint64_t r_hi = ~0ULL);
cudaError_t launch_table_fill(const TableView &t, const BinView &b, const uint32_t *skip_bits, int set_filter,
                              unsigned long long *work_counter, uint64_t n_unbinned, int num_sms, cudaStream_t s);
cudaError_t launch_search_contained(const SearchParams &p, int num_sms, cudaStream_t s);
// ev_probe_done / ev_verify_done (optional) are recorded after the probe and verify kernels of the edge pass
cudaError_t launch_search_edges(const SearchParams &p, int num_sms, cudaStream_t s, cudaEvent_t ev_probe_done = nullptr,
                                cudaEvent_t ev_verify_done = nullptr);
cudaError_t launch_contained_finish(const unsigned long long *best, uint64_t n, uint32_t *bits,
                                    unsigned long long *count, cudaStream_t s);
// rows for contained reads: keys -> (contained, container, orient, start), compacted in read order
cudaError_t launch_contained_rows(const unsigned long long *best, const ReadsView &r, int K, void *rows_out,
                                  unsigned long long *cursor, cudaStream_t s);
// out = the rows of `in` whose contained read lies in [lo, hi) (order not kept); *cursor (zeroed by the caller) = how many
cudaError_t launch_crows_in_range(const void *in, uint64_t n, uint64_t lo, uint64_t hi, void *out, unsigned long long *cursor, cudaStream_t s);
cudaError_t launch_reduce_mark(const ReduceParams &p, int num_sms, cudaStream_t s);
cudaError_t launch_reduce_emit(const ReduceParams &p, int num_sms, cudaStream_t s);
cudaError_t launch_rebase_rowinfo(uint64_t *rowinfo, uint64_t u_lo, uint64_t u_hi, uint64_t base, cudaStream_t s);
// mine[i] = min over ranks of peers[r][i] (peers = device array of the ranks' key arrays, [rank] == mine)
cudaError_t launch_min_keys(unsigned long long *mine, const uint64_t *const *peers, uint32_t world, uint32_t rank, uint64_t n, cudaStream_t s);
// sparse exchange of the containment keys: the keys that are set as (read, key) pairs / minima taken from such pairs
cudaError_t launch_compact_keys(const unsigned long long *best, uint64_t n, unsigned long long *pairs, uint64_t cap, unsigned long long *count, cudaStream_t s);
cudaError_t launch_apply_keys(unsigned long long *best, uint64_t n, const unsigned long long *pairs, uint64_t npairs, cudaStream_t s);
// out[r][0..3] = the 128 bases that end read r (bases before the read: zero)
cudaError_t launch_make_tails(const ReadsView &r, uint64_t *out, cudaStream_t s);
// out[r] = reverse complement of read r, same row layout as r.words
cudaError_t launch_revcomp_rows(const ReadsView &r, uint64_t *out, cudaStream_t s);
// copy n rows of src_words u64 (pitch src_stride) into rows of dst_stride u64, zero-filling the tail (dst_stride >= src_words)
cudaError_t launch_restride(const uint64_t *src, int src_stride, int src_words, uint64_t *dst, int dst_stride, uint64_t n, cudaStream_t s);
// smem bytes a search / reduce block needs for the given shape (0 = does not fit)
bool search_edges_fits(int max_len, int K, int cap);
// the flat edge pass (one lane per read / probe / candidate) handles this shape; otherwise the warp-per-read kernels run
bool edges_flat_supported(int max_len, int stride, int K);
// candidate-buffer entries the flat probe kernel may leave unused (one partly filled slice per resident warp)
uint64_t edges_flat_slack(int num_sms);
// simplify.cu: composite-edge contraction + dead-end removal on a reduced edge list; allocates *d_out / *d_inner (cudaFree)
// reduced edges sorted by (src, dst) on the device (simplify.cu)
cudaError_t sort_edges_device(disco_edge *d_edges, uint64_t n, uint64_t n_reads, cudaStream_t s, unsigned long long *launches);
cudaError_t run_simplify(const disco_edge *d_edges, uint64_t ne, uint64_t n_reads, const uint16_t *d_len, int uniform_len,
                         uint32_t min_ovl, uint32_t min_reads, uint32_t min_len, cudaStream_t s,
                         disco_cedge **d_out, uint64_t *n_out, uint64_t **d_inner, uint64_t *n_inner, uint64_t *rounds, uint64_t *removed_edges,
                         uint64_t *cycle_atoms, unsigned long long *launches);
void count_launches(unsigned long long k);
// kernels launched by this library since it was loaded (every launcher counts its own)
unsigned long long launches_total();

} // namespace disco
