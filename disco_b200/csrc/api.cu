// api.cu -- the C ABI of include/disco_gpu.h: context, device buffers, phase orchestration, result read-back.
// No CPU implementation of the path exists in this library: every entry point needs a CUDA device or fails.
#include "../../include/disco_gpu.h"
#include "dna.cuh"
#include "kernels.cuh"
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

using namespace disco;

namespace {
thread_local std::string g_create_error;

enum Cursor { CUR_WORK = 0, CUR_WORK2, CUR_WORK3, CUR_ROWS, CUR_EDGES, CUR_NCONTAINED, CUR_CROWS, CUR_CANDS, CUR_TABLE_FULL, CUR_BIN_OVERFLOW, CUR_COUNT };
enum Ev { EV_T0 = 0, EV_TABLE_ALL, EV_CONTAINED, EV_FINISH, EV_TABLE_NC, EV_EDGES, EV_MARK, EV_EMIT, EV_EDGES_K0, EV_EDGES_K1,
          EV_CONT_K0, EV_CONT_K1, EV_PROBE_K1, EV_VERIFY_K1, EV_MARK_K0, EV_EMIT_K0, EV_COUNT };
} // namespace

struct disco_ctx {
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err;

    // reads
    uint64_t *d_words = nullptr;
    uint16_t *d_len = nullptr;
    uint64_t *d_words_rc = nullptr; // reverse complement of every read (verify kernel: suffix overlaps read a prefix); optional
    uint64_t *d_tails = nullptr;    // last 128 bases of every read as one 32-byte sector (flat verify kernel); 64-byte rows only
    uint64_t *d_stage = nullptr; // host rows arrive here when their pitch differs from the device row
    uint64_t stage_words = 0;
    bool own_reads = true;       // false: d_words / d_len are the caller's device buffers (disco_gpu_use_reads_device)
    // deferred upload (disco_gpu_load_reads_async): host buffers waiting to be copied, chunk by chunk, under the table build
    const uint64_t *pend_packed = nullptr;
    const uint16_t *pend_len = nullptr;
    uint32_t pend_wpr = 0;
    bool pending = false;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_copy[17] = {};
    ReadsView reads{};
    // run parameters
    int K = 0, cap = 0;
    bool begun = false, have_contained = false, have_edges = false, have_reduced = false;
    bool table_has_contained = true; // the table was built from all reads (phase_table(0)) and not rebuilt since
    // table
    uint64_t *d_slots = nullptr;
    uint64_t nbuckets = 0;
    uint32_t *d_filter = nullptr;
    uint64_t filter_bits = 0;
    // binned build (kernels.cu: k_table_bin / k_table_fill): the records sorted by bucket range; kept for the rebuild
    ulonglong2 *d_bins = nullptr;
    unsigned long long *d_bin_count = nullptr;
    uint64_t bin_cap = 0;
    uint32_t nbins = 0;       // 0: direct inserts
    bool bins_valid = false;  // the bins hold this run's records
    // containment
    unsigned long long *d_best = nullptr;
    uint32_t *d_bits = nullptr;
    disco_crow *d_crows = nullptr;
    uint64_t crows_cap = 0;
    disco_crow *d_crows_part = nullptr; // disco_gpu_get_contained_range: the rows of one read range, compacted
    uint64_t crows_part_cap = 0;
    uint64_t n_contained = 0;
    uint64_t run_n = 0; // reads the run buffers are sized for
    // adjacency
    uint64_t *d_rowinfo = nullptr;
    uint64_t *d_rows = nullptr;        // owned
    uint64_t *d_rows_active = nullptr; // what the reduction reads: d_rows, or a caller-owned gathered buffer (use_rows)
    uint64_t rows_cap = 0, rows_used = 0; // rows_used = cursor (includes warp-slice slack)
    // flat edge pass: the batches' candidate lists (probe -> verify), internal to one launch
    uint64_t *d_cands = nullptr;
    uint64_t cands_cap = 0;
    uint64_t *d_batchinfo = nullptr;
    uint64_t batch_cap = 0; // batches d_batchinfo holds
    unsigned long long launches_at_begin = 0;
    // output
    disco_edge *d_edges = nullptr;
    uint64_t edges_cap = 0, n_edges = 0;
    // optional result sink: the caller's pinned host buffer, filled by the emission kernel itself (no D2H copy afterwards)
    disco_edge *sink_host = nullptr, *sink_dev = nullptr;
    uint64_t sink_cap = 0;
    bool sink_filled = false;
    // simplified graph (disco_gpu_simplify)
    disco_cedge *d_cedges = nullptr;
    uint64_t *d_inner = nullptr;
    uint64_t n_cedges = 0, n_inner = 0, simp_rounds = 0, simp_removed = 0, simp_cycle = 0;
    float simp_ms = 0.f;
    // counters / stats
    unsigned long long *d_cursors = nullptr;  // CUR_COUNT
    unsigned long long *d_stats_c = nullptr;  // containment pass
    unsigned long long *d_stats_e = nullptr;  // edge pass + reduction
    float acc_probe = 0.f, acc_verify = 0.f, acc_exact = 0.f, acc_edges = 0.f; // edge-pass kernel times summed over parts
    // key-sharded mode (Mode B): this context holds shard `shard_rank` of `shard_world` of the table and the adjacency
    // rows of its own query range; the other shards are reached through peer-mapped pointers (CUDA IPC)
    bool own_slots = true, own_rows = true; // false: caller-owned memory (disco_gpu_adopt_buffer), never freed or grown here
    uint32_t shard_world = 1, shard_rank = 0;
    bool table_sharded = true; // false: the table is replicated (built from all reads on this GPU); only the adjacency is partitioned
    uint32_t tw() const { return table_sharded ? shard_world : 1u; } // shards of the table
    struct PeerSet {
        const uint64_t **d_ptrs = nullptr;         // device array [DISCO_MAX_SHARDS]
        void *opened[DISCO_MAX_SHARDS] = {};       // mappings this context opened (to close them again)
        cudaIpcMemHandle_t handle[DISCO_MAX_SHARDS] = {};
        bool ready = false;
    } peer_table, peer_rows, peer_keys;
    uint64_t *d_bounds = nullptr; // read-id bounds of the ranks' query ranges [shard_world + 1]
    cudaEvent_t ev[EV_COUNT] = {};
    bool ev_done[EV_COUNT] = {};
    disco_stats stats{};
};

namespace {

int fail(disco_ctx *c, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf; else g_create_error = buf;
    return code;
}

#define CK(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t e__ = (call);                                                                             \
        if (e__ != cudaSuccess)                                                                               \
            return fail(ctx, e__ == cudaErrorMemoryAllocation ? DISCO_E_NOMEM : DISCO_E_CUDA, "%s failed: %s (%s:%d)", #call, \
                        cudaGetErrorString(e__), __FILE__, __LINE__);                                         \
    } while (0)

template <typename T>
void dfree(T *&p)
{
    if (p) cudaFree(p);
    p = nullptr;
}

// first guess of adjacency entries / candidates per query read (30x, 150 bp, minOverlap 50 needs ~37); a pass that
// overflows is repeated with the exact size.  DISCO_ENTRIES_PER_READ: tests force the retry path with a tiny guess.
uint64_t entries_per_read_guess()
{
    const char *e = getenv("DISCO_ENTRIES_PER_READ");
    const long v = e ? atol(e) : 48;
    return v >= 1 && v <= 4096 ? (uint64_t)v : 48;
}

int pick_stride(int max_len)
{
    const int W = (max_len + 31) / 32;
    // power-of-two rows (16..128 bytes) never straddle a 64-byte DRAM fetch / 128-byte line: one random access per
    // candidate read
    const int regs[] = {2, 4, 8, 16};
    for (int s : regs) if (W <= s) return s;
    return (W + 1) & ~1; // long reads: generic (global-walking) matcher, still 16-byte aligned rows
}

void free_run_buffers(disco_ctx *c)
{
    if (c->own_slots) dfree(c->d_slots);
    if (c->own_rows) dfree(c->d_rows);
    c->d_slots = nullptr; c->d_rows = nullptr;
    c->own_slots = c->own_rows = true;
    c->peer_table.ready = c->peer_rows.ready = false; // whatever the peers mapped is gone
    dfree(c->d_filter); dfree(c->d_best); dfree(c->d_bits); dfree(c->d_crows); dfree(c->d_crows_part); c->crows_part_cap = 0;
    dfree(c->d_bins); dfree(c->d_bin_count); c->nbins = 0; c->bin_cap = 0; c->bins_valid = false;
    dfree(c->d_rowinfo); dfree(c->d_edges); dfree(c->d_cands); dfree(c->d_batchinfo); dfree(c->d_cedges); dfree(c->d_inner);
    c->n_cedges = c->n_inner = 0;
    c->d_rows_active = nullptr;
    c->rows_cap = c->edges_cap = c->crows_cap = c->run_n = c->cands_cap = c->batch_cap = 0;
    c->begun = c->have_contained = c->have_edges = c->have_reduced = false;
}

void free_reads(disco_ctx *c)
{
    if (c->own_reads) { dfree(c->d_words); dfree(c->d_len); }
    c->d_words = nullptr; c->d_len = nullptr; c->own_reads = true;
    dfree(c->d_words_rc); dfree(c->d_tails); dfree(c->d_stage);
    c->stage_words = 0;
    c->pending = false; c->pend_packed = nullptr; c->pend_len = nullptr;
    c->reads = ReadsView{};
}

void close_peers(disco_ctx::PeerSet &ps)
{
    for (int r = 0; r < DISCO_MAX_SHARDS; r++) {
        if (ps.opened[r]) cudaIpcCloseMemHandle(ps.opened[r]);
        ps.opened[r] = nullptr;
    }
    ps.ready = false;
}

// candidate lists of the flat kernels for a launch over np reads: four segment words per 32-read batch + the lists
int ensure_flat_buffers(disco_ctx *ctx, uint64_t np)
{
    const uint64_t nb = (np + 31) / 32;
    if (nb > ctx->batch_cap) {
        dfree(ctx->d_batchinfo);
        ctx->batch_cap = 0;
        CK(cudaMalloc(&ctx->d_batchinfo, std::max<uint64_t>(nb, 1) * 4 * sizeof(uint64_t)));
        ctx->batch_cap = nb;
    }
    const bool tiny = getenv("DISCO_ENTRIES_PER_READ") != nullptr;
    const uint64_t slack = tiny ? 8192 : std::min<uint64_t>(edges_flat_slack(ctx->num_sms), nb * 8192) + 8192;
    const uint64_t want = std::max<uint64_t>((np + 1024) * entries_per_read_guess(), tiny ? 1 << 10 : 1 << 16) + slack;
    if (ctx->cands_cap < want) { // (a buffer a retry has grown is kept; parts of one pass differ by a read at most)
        size_t free_b = 0, total_b = 0;
        CK(cudaMemGetInfo(&free_b, &total_b));
        const uint64_t lim = (uint64_t)((free_b + ctx->cands_cap * sizeof(uint64_t)) * 0.5) / sizeof(uint64_t);
        const uint64_t take = std::min(want, std::max<uint64_t>(lim, 1 << 16));
        if (take > ctx->cands_cap) {
            dfree(ctx->d_cands);
            ctx->cands_cap = 0;
            CK(cudaMalloc(&ctx->d_cands, take * sizeof(uint64_t)));
            ctx->cands_cap = take;
        }
    }
    return DISCO_OK;
}

TableView table_view(const disco_ctx *c)
{
    TableView tv{};
    tv.slots = c->d_slots; tv.nbuckets = c->nbuckets; tv.filter = c->d_filter;
    tv.filter_mask = (uint32_t)(c->filter_bits ? c->filter_bits - 1 : 0);
    tv.peers = c->peer_table.d_ptrs; tv.world = c->tw(); tv.rank = c->table_sharded ? c->shard_rank : 0u;
    tv.full = reinterpret_cast<unsigned int *>(c->d_cursors + CUR_TABLE_FULL);
    return tv;
}

int record(disco_ctx *ctx, int which)
{
    CK(cudaEventRecord(ctx->ev[which], ctx->stream));
    ctx->ev_done[which] = true;
    return DISCO_OK;
}

int alloc_reads(disco_ctx *ctx, uint64_t n, int min_len, int max_len)
{
    // same shape as the resident set (the steady state of a service that processes batch after batch): keep every buffer
    ctx->pending = false;
    if (ctx->d_words && ctx->own_reads && n == ctx->reads.n && n > 0 && pick_stride(max_len) == ctx->reads.stride) {
        ctx->reads.min_len = min_len; ctx->reads.max_len = max_len;
        ctx->reads.uniform_len = (min_len == max_len) ? max_len : 0;
        ctx->begun = ctx->have_contained = ctx->have_edges = ctx->have_reduced = false;
        return DISCO_OK;
    }
    free_run_buffers(ctx);
    free_reads(ctx);
    if (n == 0) return fail(ctx, DISCO_E_ARG, "no reads");
    if (n > 0x7FFFFFF0ULL) return fail(ctx, DISCO_E_LIMIT, "at most 2^31-16 reads per context (record ids are 32 bit)");
    if (max_len > 32767) return fail(ctx, DISCO_E_LIMIT, "read length %d exceeds the 15-bit limit of the reference record header (HashTable.cpp:437)", max_len);
    if (min_len < 1) return fail(ctx, DISCO_E_ARG, "empty read");
    const int stride = pick_stride(max_len);
    CK(cudaMalloc(&ctx->d_words, n * (uint64_t)stride * sizeof(uint64_t)));
    CK(cudaMalloc(&ctx->d_len, n * sizeof(uint16_t)));
    // Optional second copy, reverse-complemented (DISCO_RC_COPY=1; rows of 32..128 bytes): the verify kernel then reads a
    // candidate's overlapping suffix as the leading sector(s) of this copy instead of the whole 64-byte row.  Measured on
    // B200, 10M x 150 bp: verify 16.5 -> 15.3 ms on the single genome, but slower on the metagenome shape and 64 bytes
    // per read more memory -- off by default.
    if (stride >= 4 && stride <= 16 && getenv("DISCO_RC_COPY") && atoi(getenv("DISCO_RC_COPY")) == 1) {
        if (cudaMalloc(&ctx->d_words_rc, n * (uint64_t)stride * sizeof(uint64_t)) != cudaSuccess) { ctx->d_words_rc = nullptr; cudaGetLastError(); }
    }
    // Tail-sector copy (32 bytes per read, rows of 64 bytes): lets the verify kernel fetch one 32-byte sector per candidate
    // for overlaps of up to 128 bases.  Optional: skipped when memory is short or DISCO_TAILS=0.
    if (stride == 8 && !(getenv("DISCO_TAILS") && atoi(getenv("DISCO_TAILS")) == 0)) {
        if (cudaMalloc(&ctx->d_tails, n * 4 * sizeof(uint64_t)) != cudaSuccess) { ctx->d_tails = nullptr; cudaGetLastError(); }
    }
    ctx->reads.tails = nullptr; // set by disco_gpu_begin once the copy is filled
    ctx->reads.words = ctx->d_words; ctx->reads.words_rc = ctx->d_words_rc; ctx->reads.len = ctx->d_len; ctx->reads.n = n; ctx->reads.stride = stride;
    ctx->reads.min_len = min_len; ctx->reads.max_len = max_len;
    ctx->reads.uniform_len = (min_len == max_len) ? max_len : 0;
    return DISCO_OK;
}

int copy_reads(disco_ctx *ctx, const uint64_t *packed, const uint16_t *len, uint32_t wpr, cudaMemcpyKind kind)
{
    const uint64_t n = ctx->reads.n;
    const int stride = ctx->reads.stride;
    const int W = (ctx->reads.max_len + 31) / 32;
    if ((int)wpr < W) return fail(ctx, DISCO_E_ARG, "words_per_read %u too small for max length %d", wpr, ctx->reads.max_len);
    if ((int)wpr == stride) { // same pitch on both sides: one flat copy
        CK(cudaMemcpyAsync(ctx->d_words, packed, n * (uint64_t)stride * sizeof(uint64_t), kind, ctx->stream));
    } else if (kind == cudaMemcpyDeviceToDevice) {
        CK(launch_restride(packed, (int)wpr, std::min<int>((int)wpr, stride), ctx->d_words, stride, n, ctx->stream));
    } else {
        // host rows with another pitch: one flat copy into a staging buffer (a pitched copy of millions of 40-byte rows
        // is far slower than PCIe), then a re-stride kernel
        const uint64_t need = n * (uint64_t)wpr;
        if (ctx->stage_words < need) {
            dfree(ctx->d_stage);
            ctx->stage_words = 0;
            CK(cudaMalloc(&ctx->d_stage, need * sizeof(uint64_t)));
            ctx->stage_words = need;
        }
        CK(cudaMemcpyAsync(ctx->d_stage, packed, need * sizeof(uint64_t), kind, ctx->stream));
        CK(launch_restride(ctx->d_stage, (int)wpr, std::min<int>((int)wpr, stride), ctx->d_words, stride, n, ctx->stream));
    }
    CK(cudaMemcpyAsync(ctx->d_len, len, n * sizeof(uint16_t), kind, ctx->stream));
    return DISCO_OK;
}

} // namespace

// ===================================================================================================================
extern "C" {

int disco_gpu_create(disco_ctx **out, int device)
{
    disco_ctx *ctx = nullptr;
    if (!out) return fail(nullptr, DISCO_E_ARG, "out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, DISCO_E_CUDA, "no CUDA device available (%s); this library has no CPU path", cudaGetErrorString(e));
    if (device < 0 || device >= count) return fail(nullptr, DISCO_E_ARG, "device %d out of range (%d devices)", device, count);
    ctx = new (std::nothrow) disco_ctx();
    if (!ctx) return fail(nullptr, DISCO_E_NOMEM, "out of host memory");
    ctx->device = device;
    auto bail = [&](cudaError_t ce, const char *what) {
        fail(nullptr, DISCO_E_CUDA, "%s failed: %s", what, cudaGetErrorString(ce));
        delete ctx;
        return DISCO_E_CUDA;
    };
    if ((e = cudaSetDevice(device)) != cudaSuccess) return bail(e, "cudaSetDevice");
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return bail(e, "cudaGetDeviceProperties");
    if (prop.major < 10) { fail(nullptr, DISCO_E_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor); delete ctx; return DISCO_E_CUDA; }
    ctx->num_sms = prop.multiProcessorCount;
    // the path is random 32-byte sectors: do not let L2 pull in the neighbouring sectors of every miss
    cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 32);
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
    ctx->own_stream = true;
    for (auto &ev : ctx->ev) if ((e = cudaEventCreate(&ev)) != cudaSuccess) return bail(e, "cudaEventCreate");
    if ((e = cudaMalloc(&ctx->d_cursors, CUR_COUNT * sizeof(unsigned long long))) != cudaSuccess) return bail(e, "cudaMalloc");
    if ((e = cudaMalloc(&ctx->d_stats_c, ST_COUNT * sizeof(unsigned long long))) != cudaSuccess) return bail(e, "cudaMalloc");
    if ((e = cudaMalloc(&ctx->d_stats_e, ST_COUNT * sizeof(unsigned long long))) != cudaSuccess) return bail(e, "cudaMalloc");
    *out = ctx;
    return DISCO_OK;
}

void disco_gpu_destroy(disco_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    close_peers(ctx->peer_table); close_peers(ctx->peer_rows);
    dfree(ctx->peer_table.d_ptrs); dfree(ctx->peer_rows.d_ptrs); dfree(ctx->peer_keys.d_ptrs); dfree(ctx->d_bounds);
    free_run_buffers(ctx);
    free_reads(ctx);
    dfree(ctx->d_cursors); dfree(ctx->d_stats_c); dfree(ctx->d_stats_e);
    for (auto &ev : ctx->ev) if (ev) cudaEventDestroy(ev);
    for (auto &ev : ctx->ev_copy) if (ev) cudaEventDestroy(ev);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int disco_gpu_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char *disco_gpu_last_error(const disco_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int disco_gpu_set_stream(disco_ctx *ctx, void *cuda_stream)
{
    if (!ctx) return DISCO_E_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream) { cudaStreamDestroy(ctx->stream); ctx->own_stream = false; }
    ctx->stream = reinterpret_cast<cudaStream_t>(cuda_stream);
    return DISCO_OK;
}

int disco_gpu_load_reads(disco_ctx *ctx, const uint64_t *packed, const uint16_t *len, uint64_t n_reads, uint32_t words_per_read)
{
    if (!ctx) return DISCO_E_ARG;
    if (!packed || !len) return fail(ctx, DISCO_E_ARG, "NULL input");
    CK(cudaSetDevice(ctx->device));
    int mn = 1 << 30, mx = 0;
    for (uint64_t i = 0; i < n_reads; i++) { int l = len[i]; mn = std::min(mn, l); mx = std::max(mx, l); }
    int rc = alloc_reads(ctx, n_reads, mn, mx);
    if (rc) return rc;
    return copy_reads(ctx, packed, len, words_per_read, cudaMemcpyHostToDevice);
}

int disco_gpu_load_reads_device(disco_ctx *ctx, const uint64_t *d_packed, const uint16_t *d_len, uint64_t n_reads,
                                uint32_t words_per_read, uint32_t min_len, uint32_t max_len)
{
    if (!ctx) return DISCO_E_ARG;
    if (!d_packed || !d_len) return fail(ctx, DISCO_E_ARG, "NULL input");
    if (min_len > max_len) return fail(ctx, DISCO_E_ARG, "min_len > max_len");
    CK(cudaSetDevice(ctx->device));
    int rc = alloc_reads(ctx, n_reads, (int)min_len, (int)max_len);
    if (rc) return rc;
    return copy_reads(ctx, d_packed, d_len, words_per_read, cudaMemcpyDeviceToDevice);
}

// The caller's device buffers used in place (no copy): rows must already have this library's pitch -- the power of two
// {2,4,8,16} that holds the longest read, else an even word count -- and 16-byte alignment; any other pitch is copied as
// disco_gpu_load_reads_device does.  The buffers must stay valid and unchanged until the results have been read.
int disco_gpu_use_reads_device(disco_ctx *ctx, const uint64_t *d_packed, const uint16_t *d_len, uint64_t n_reads,
                               uint32_t words_per_read, uint32_t min_len, uint32_t max_len)
{
    if (!ctx) return DISCO_E_ARG;
    if (!d_packed || !d_len) return fail(ctx, DISCO_E_ARG, "NULL input");
    if (min_len > max_len || min_len < 1) return fail(ctx, DISCO_E_ARG, "bad length bounds");
    if (max_len > 32767) return fail(ctx, DISCO_E_LIMIT, "read length %u exceeds the 15-bit limit of the reference record header (HashTable.cpp:437)", max_len);
    const int stride = pick_stride((int)max_len);
    if ((int)words_per_read != stride || (reinterpret_cast<uintptr_t>(d_packed) & 15))
        return disco_gpu_load_reads_device(ctx, d_packed, d_len, n_reads, words_per_read, min_len, max_len);
    if (n_reads == 0) return fail(ctx, DISCO_E_ARG, "no reads");
    if (n_reads > 0x7FFFFFF0ULL) return fail(ctx, DISCO_E_LIMIT, "at most 2^31-16 reads per context (record ids are 32 bit)");
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    const bool same = ctx->d_words && n_reads == ctx->reads.n && stride == ctx->reads.stride;
    if (same) { // keep the run buffers (and the tail copy's allocation); only the read buffers change hands
        if (ctx->own_reads) { dfree(ctx->d_words); dfree(ctx->d_len); }
        dfree(ctx->d_words_rc);
        dfree(ctx->d_stage); ctx->stage_words = 0;
    } else {
        free_run_buffers(ctx);
        free_reads(ctx);
        if (stride == 8 && !(getenv("DISCO_TAILS") && atoi(getenv("DISCO_TAILS")) == 0)) {
            if (cudaMalloc(&ctx->d_tails, n_reads * 4 * sizeof(uint64_t)) != cudaSuccess) { ctx->d_tails = nullptr; cudaGetLastError(); }
        }
    }
    ctx->own_reads = false; ctx->pending = false;
    ctx->d_words = const_cast<uint64_t *>(d_packed); ctx->d_len = const_cast<uint16_t *>(d_len);
    ctx->reads.tails = nullptr;
    ctx->reads.words = ctx->d_words; ctx->reads.words_rc = nullptr; ctx->reads.len = ctx->d_len; ctx->reads.n = n_reads; ctx->reads.stride = stride;
    ctx->reads.min_len = (int)min_len; ctx->reads.max_len = (int)max_len;
    ctx->reads.uniform_len = (min_len == max_len) ? (int)max_len : 0;
    ctx->begun = ctx->have_contained = ctx->have_edges = ctx->have_reduced = false;
    return DISCO_OK;
}

// disco_gpu_load_reads with the copy deferred into disco_gpu_build_graph / disco_gpu_phase_table(ctx, 0): there the rows
// cross PCIe in chunks on a copy stream while the table is built from the chunks that have arrived, so the table build
// (and the re-striding) hide under the transfer.  The host buffers must stay valid and unchanged until that call has
// returned; only page-locked memory makes the copies asynchronous.  min_len / max_len: the shortest and longest read --
// every loader knows them (Dataset.cpp prints them) -- or 0, 0 to have them found here (one pass over `len`).
int disco_gpu_load_reads_async(disco_ctx *ctx, const uint64_t *packed, const uint16_t *len, uint64_t n_reads, uint32_t words_per_read,
                               uint32_t min_len, uint32_t max_len)
{
    if (!ctx) return DISCO_E_ARG;
    if (!packed || !len) return fail(ctx, DISCO_E_ARG, "NULL input");
    CK(cudaSetDevice(ctx->device));
    int mn = (int)min_len, mx = (int)max_len;
    if (!mn || !mx) {
        mn = 1 << 30; mx = 0;
        for (uint64_t i = 0; i < n_reads; i++) { int l = len[i]; mn = std::min(mn, l); mx = std::max(mx, l); }
    }
    if (mn > mx) return fail(ctx, DISCO_E_ARG, "min_len > max_len");
    int rc = alloc_reads(ctx, n_reads, mn, mx);
    if (rc) return rc;
    if ((int)words_per_read < (mx + 31) / 32) return fail(ctx, DISCO_E_ARG, "words_per_read %u too small for max length %d", words_per_read, mx);
    if ((int)words_per_read != ctx->reads.stride) {
        const uint64_t need = n_reads * (uint64_t)words_per_read;
        if (ctx->stage_words < need) {
            dfree(ctx->d_stage);
            ctx->stage_words = 0;
            CK(cudaMalloc(&ctx->d_stage, need * sizeof(uint64_t)));
            ctx->stage_words = need;
        }
    }
    if (!ctx->copy_stream) {
        CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        for (auto &e : ctx->ev_copy) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    ctx->pend_packed = packed; ctx->pend_len = len; ctx->pend_wpr = words_per_read;
    ctx->pending = true;
    return DISCO_OK;
}

// ---- phases -------------------------------------------------------------------------------------------------------
int disco_gpu_begin(disco_ctx *ctx, uint32_t min_overlap, uint32_t max_edge_per_kmer)
{
    if (!ctx) return DISCO_E_ARG;
    if (!ctx->d_words) return fail(ctx, DISCO_E_ARG, "load reads first");
    if (min_overlap < 2) return fail(ctx, DISCO_E_ARG, "min_overlap must be >= 2");
    if (max_edge_per_kmer < 1 || max_edge_per_kmer > 8) return fail(ctx, DISCO_E_ARG, "max_edge_per_kmer must be in 1..8");
    if ((uint32_t)ctx->reads.min_len <= min_overlap)
        return fail(ctx, DISCO_E_ARG, "every read must be longer than min_overlap (Dataset.cpp:305): shortest is %d", ctx->reads.min_len);
    CK(cudaSetDevice(ctx->device));
    ctx->begun = ctx->have_contained = ctx->have_edges = ctx->have_reduced = false;
    ctx->K = (int)min_overlap - 1; // hashStringLength (HashTable.cpp:50)
    ctx->cap = (int)max_edge_per_kmer;
    if (!search_edges_fits(ctx->reads.max_len, ctx->K, ctx->cap))
        return fail(ctx, DISCO_E_LIMIT, "max read length %d with min_overlap %u needs more shared memory per warp than one SM has", ctx->reads.max_len, min_overlap);
    const uint64_t n = ctx->reads.n;
    // 2n records, four 8-byte slots per 32-byte bucket (the reference sizes its table at 8n+1 index words,
    // HashTable.cpp:53)
    if (ctx->run_n != n) {
        free_run_buffers(ctx);
        {   // load factor 1/6 (3n buckets) keeps ~99.5% of the look-ups to a single 32-byte sector; fall back to 1/3 when
            // that would take more than a tenth of the free memory
            size_t free_b = 0, total_b = 0;
            CK(cudaMemGetInfo(&free_b, &total_b));
            uint64_t nb = 3 * n;
            if (const char *e = getenv("DISCO_TABLE_BUCKETS_X10")) nb = n * (uint64_t)std::max(6, std::min(80, atoi(e))) / 10; // (tuning knob)
            if (nb * 32 > free_b / 10 * ctx->tw()) nb = n + n / 2;
            nb = (nb + ctx->tw() - 1) / ctx->tw(); // key-sharded: buckets of this GPU's shard
            ctx->nbuckets = std::max<uint64_t>(1024, nb);
        }
        CK(cudaMalloc(&ctx->d_slots, ctx->nbuckets * 4 * sizeof(uint64_t)));
        // presence filter: 16 bits per record, at most 32 MB: measured (profiles/filter_bench.cu, profiles/README.md) the
        // look-ups run at full L2 speed up to 16 MB, -5% at 32 MB, -17% at 64 MB while 1.2 GB of buckets stream by; at
        // 10 M reads 32 MB is the best trade between look-up speed and false-positive bucket reads.  Pointless once it
        // has fewer bits than records.
        // (a replicated table over 80 M reads, one rank's 10 M queries -- profiles/probe_big_table.py: 2^28 bits probe 18.7 ms /
        //  table 11.7 ms, 2^29: 18.6 / 12.7, 2^30: 20.8 / 14.6, 2^31: 24.7 / 16.0 -- the false positives a larger filter saves
        //  cost less than its own misses: stay L2 resident)
        const uint64_t filter_cap = n > 160000000ULL ? (1ULL << 29) : (1ULL << 28);
        ctx->filter_bits = 1ULL << 16;
        while (ctx->filter_bits < 32 * n && ctx->filter_bits < filter_cap) ctx->filter_bits <<= 1;
        if (const char *e = getenv("DISCO_FILTER_LOG2")) { // tuning knob: 0 disables the filter
            const int lg = atoi(e);
            ctx->filter_bits = lg >= 10 && lg <= 33 ? (1ULL << lg) : 0;
        }
        if (ctx->filter_bits < 2 * n) ctx->filter_bits = 0;
        if (ctx->filter_bits) CK(cudaMalloc(&ctx->d_filter, ctx->filter_bits / 8));
        CK(cudaMalloc(&ctx->d_best, n * sizeof(unsigned long long)));
        CK(cudaMalloc(&ctx->d_bits, ((n + 31) / 32) * sizeof(uint32_t)));
        CK(cudaMalloc(&ctx->d_rowinfo, n * sizeof(uint64_t)));
        // Binned build for tables that do not fit L2 (single / replicated tables, short reads): one bin per 16 MB of
        // table, 25 % head room per bin (a bin that overflows -- skewed k-mers -- hands the build to the direct kernel).
        // DISCO_BINNED=0 keeps the direct inserts.
        ctx->nbins = 0; ctx->bin_cap = 0;
        if (ctx->tw() == 1 && table_bin_supported(ctx->reads.max_len) && !(getenv("DISCO_BINNED") && atoi(getenv("DISCO_BINNED")) == 0)) {
            uint64_t slice = 16ULL << 20;
            if (const char *e = getenv("DISCO_BIN_SLICE_KB")) slice = (uint64_t)std::max(1, atoi(e)) << 10; // (tests: bins on small inputs)
            uint32_t nb = 1;
            while ((uint64_t)nb * slice < ctx->nbuckets * 32 && nb < 1024) nb <<= 1;
            const bool unbinned = getenv("DISCO_UNBINNED") && atoi(getenv("DISCO_UNBINNED")) == 1; // records in read order (A/B)
            if (unbinned && nb >= 4) nb = 1;
            if (nb >= 4 || unbinned) {
                uint64_t cap = unbinned ? 2 * n : (2 * n / nb) + (2 * n / nb) / 4 + 4096;
                if (const char *e = getenv("DISCO_BIN_HEADROOM")) cap = (2 * n / nb) + (uint64_t)std::max(0, atoi(e)); // (tests: force the overflow path)
                size_t free_b = 0, total_b = 0;
                CK(cudaMemGetInfo(&free_b, &total_b));
                if ((uint64_t)nb * cap * sizeof(ulonglong2) < free_b / 4 &&
                    cudaMalloc(&ctx->d_bins, (uint64_t)nb * cap * sizeof(ulonglong2)) == cudaSuccess) {
                    CK(cudaMalloc(&ctx->d_bin_count, nb * sizeof(unsigned long long)));
                    ctx->nbins = nb; ctx->bin_cap = cap;
                } else {
                    cudaGetLastError();
                }
            }
        }
        ctx->run_n = n;
    }
    ctx->bins_valid = false;
    if (ctx->pending) CK(cudaEventRecord(ctx->ev_copy[16], ctx->stream));
    CK(cudaMemsetAsync(ctx->d_best, 0xFF, n * sizeof(unsigned long long), ctx->stream));
    CK(cudaMemsetAsync(ctx->d_bits, 0, ((n + 31) / 32) * sizeof(uint32_t), ctx->stream));
    CK(cudaMemsetAsync(ctx->d_rowinfo, 0, n * sizeof(uint64_t), ctx->stream));
    CK(cudaMemsetAsync(ctx->d_cursors, 0, CUR_COUNT * sizeof(unsigned long long), ctx->stream));
    CK(cudaMemsetAsync(ctx->d_stats_c, 0, ST_COUNT * sizeof(unsigned long long), ctx->stream));
    CK(cudaMemsetAsync(ctx->d_stats_e, 0, ST_COUNT * sizeof(unsigned long long), ctx->stream));
    ctx->stats = disco_stats{};
    ctx->launches_at_begin = launches_total();
    ctx->stats.n_reads = n;
    ctx->stats.table_buckets = ctx->nbuckets * ctx->tw();
    for (auto &d : ctx->ev_done) d = false;
    ctx->n_contained = ctx->n_edges = 0;
    ctx->begun = true;
    int rc0 = record(ctx, EV_T0);
    if (rc0) return rc0;
    // (a deferred upload prepares these copies chunk by chunk in disco_gpu_phase_table)
    if (ctx->d_words_rc && !ctx->pending) CK(launch_revcomp_rows(ctx->reads, ctx->d_words_rc, ctx->stream)); // part of the timed run
    ctx->reads.tails = nullptr;
    if (ctx->d_tails && ctx->reads.uniform_len > 128) { // (every read one length: the kernel knows the overlap before it fetches)
        if (!ctx->pending) CK(launch_make_tails(ctx->reads, ctx->d_tails, ctx->stream));      // part of the timed run
        ctx->reads.tails = ctx->d_tails;
    }
    return DISCO_OK;
}

namespace {
BinView bin_view(const disco_ctx *c)
{
    BinView b{};
    b.recs = c->d_bins; b.count = c->d_bin_count; b.cap = c->bin_cap; b.nbins = c->nbins;
    b.overflow = reinterpret_cast<unsigned int *>(c->d_cursors + CUR_BIN_OVERFLOW);
    return b;
}

// The deferred upload (disco_gpu_load_reads_async) and the first table build as one pipeline: chunk c crosses PCIe on the
// copy stream while the main stream re-strides chunk c-1, derives its tail sectors and inserts its records.
int upload_and_insert(disco_ctx *ctx, const TableView &tv)
{
    const uint64_t n = ctx->reads.n;
    const int stride = ctx->reads.stride;
    const uint32_t wpr = ctx->pend_wpr;
    const bool flat = (int)wpr == stride;
    int chunks = 8;
    if (const char *e = getenv("DISCO_UPLOAD_CHUNKS")) chunks = std::max(1, std::min(16, atoi(e)));
    if (n < 65536) chunks = 1;
    // whatever still reads the previous batch's rows was queued before disco_gpu_begin: the copies start after that point
    // (ev_copy[16], recorded there) and do not wait for this run's memsets
    CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_copy[16], 0));
    if (!ctx->reads.uniform_len) // (one length: the kernels never look at the array)
        CK(cudaMemcpyAsync(ctx->d_len, ctx->pend_len, n * sizeof(uint16_t), cudaMemcpyHostToDevice, ctx->copy_stream));
    for (int c = 0; c < chunks; c++) {
        const uint64_t lo = n * (uint64_t)c / chunks, hi = n * (uint64_t)(c + 1) / chunks, m = hi - lo;
        if (!m) continue;
        uint64_t *dst = flat ? ctx->d_words + lo * stride : ctx->d_stage + lo * wpr;
        CK(cudaMemcpyAsync(dst, ctx->pend_packed + lo * wpr, m * wpr * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->copy_stream));
        CK(cudaEventRecord(ctx->ev_copy[c], ctx->copy_stream));
        CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_copy[c], 0));
        if (!flat) CK(launch_restride(ctx->d_stage + lo * wpr, (int)wpr, std::min<int>((int)wpr, stride), ctx->d_words + lo * stride, stride, m, ctx->stream));
        ReadsView part = ctx->reads;
        part.words += lo * stride; part.len += lo; part.n = m; part.tails = nullptr;
        if (ctx->reads.tails) CK(launch_make_tails(part, ctx->d_tails + lo * 4, ctx->stream));
        if (ctx->d_words_rc) CK(launch_revcomp_rows(part, ctx->d_words_rc + lo * stride, ctx->stream));
        if (ctx->nbins) CK(launch_table_bin(ctx->reads, tv, ctx->K, bin_view(ctx), ctx->num_sms, ctx->stream, lo, hi));
        else CK(launch_table_insert(ctx->reads, tv, ctx->K, nullptr, ctx->num_sms, ctx->stream, lo, hi));
    }
    ctx->pending = false;
    return DISCO_OK;
}
} // namespace

int disco_gpu_phase_table(disco_ctx *ctx, int exclude_contained)
{
    if (!ctx || !ctx->begun) return fail(ctx, DISCO_E_ARG, "call disco_gpu_begin first");
    if (exclude_contained && !ctx->have_contained) return fail(ctx, DISCO_E_ARG, "containment pass not finished");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemsetAsync(ctx->d_slots, 0xFF, ctx->nbuckets * 4 * sizeof(uint64_t), ctx->stream));
    // (rebuild from the bins: the presence filter of the first build is kept -- a superset: the k-mers only contained reads
    //  have cost a wasted bucket read when probed; clearing and re-setting 2n bits costs more -- 20 M-read table: 2.39 -> 1.52 ms)
    const bool keep_filter = exclude_contained && ctx->nbins && ctx->bins_valid;
    if (ctx->d_filter && !keep_filter) CK(cudaMemsetAsync(ctx->d_filter, 0, ctx->filter_bits / 8, ctx->stream));
    ctx->table_has_contained = !exclude_contained;
    const TableView tv = table_view(ctx);
    const uint32_t *skip = exclude_contained ? ctx->d_bits : nullptr;
    if (ctx->pending && exclude_contained) return fail(ctx, DISCO_E_ARG, "reads not uploaded yet");
    if (!ctx->nbins) { // direct inserts
        if (ctx->pending) { const int rc = upload_and_insert(ctx, tv); if (rc) return rc; }
        else CK(launch_table_insert(ctx->reads, tv, ctx->K, skip, ctx->num_sms, ctx->stream));
        return record(ctx, exclude_contained ? EV_TABLE_NC : EV_TABLE_ALL);
    }
    // binned build: hash + bin (first build of a run; the rebuild without the contained reads re-uses the bins), fill bin by
    // bin; if a bin overflowed (flag on the device, no host round trip) the fill does nothing and the gated direct kernel runs
    const BinView bv = bin_view(ctx);
    if (!ctx->bins_valid) {
        CK(cudaMemsetAsync(ctx->d_bin_count, 0, ctx->nbins * sizeof(unsigned long long), ctx->stream));
        CK(cudaMemsetAsync(ctx->d_cursors + CUR_BIN_OVERFLOW, 0, sizeof(unsigned long long), ctx->stream));
        if (ctx->pending) { const int rc = upload_and_insert(ctx, tv); if (rc) return rc; }
        else CK(launch_table_bin(ctx->reads, tv, ctx->K, bv, ctx->num_sms, ctx->stream));
        ctx->bins_valid = true;
        CK(cudaMemsetAsync(ctx->d_cursors + CUR_WORK, 0, sizeof(unsigned long long), ctx->stream));
        CK(launch_table_fill(tv, bv, skip, skip ? 1 : 0, ctx->d_cursors + CUR_WORK, 2 * ctx->reads.n, ctx->num_sms, ctx->stream));
    } else {
        CK(cudaMemsetAsync(ctx->d_cursors + CUR_WORK, 0, sizeof(unsigned long long), ctx->stream));
        CK(launch_table_fill(tv, bv, skip, 0, ctx->d_cursors + CUR_WORK, 2 * ctx->reads.n, ctx->num_sms, ctx->stream));
    }
    CK(launch_table_insert(ctx->reads, tv, ctx->K, skip, ctx->num_sms, ctx->stream, 0, ~0ULL, bv.overflow));
    return record(ctx, exclude_contained ? EV_TABLE_NC : EV_TABLE_ALL);
}

int disco_gpu_phase_contained(disco_ctx *ctx, uint64_t q_lo, uint64_t q_hi)
{
    if (!ctx || !ctx->begun) return fail(ctx, DISCO_E_ARG, "call disco_gpu_begin first");
    if (q_lo > q_hi || q_hi > ctx->reads.n) return fail(ctx, DISCO_E_ARG, "bad query range");
    if (ctx->tw() > 1 && !ctx->peer_table.ready) return fail(ctx, DISCO_E_ARG, "sharded table: import the peers' shards first (disco_gpu_import_peers)");
    CK(cudaSetDevice(ctx->device));
    SearchParams p{};
    p.reads = ctx->reads; p.table = table_view(ctx);
    p.K = ctx->K; p.cap = ctx->cap; p.q_lo = q_lo; p.q_hi = q_hi;
    p.work_counter = ctx->d_cursors + CUR_WORK; p.stats = ctx->d_stats_c; p.best = ctx->d_best;
    p.rowinfo = ctx->d_rowinfo;
    // reads of several lengths: the flat kernels (candidate lists in global memory; a list that does not fit is measured by
    // the cursor and the pass repeated).  One length: k_contain_uniform, no lists.
    const bool flat = !ctx->reads.uniform_len && q_hi > q_lo && edges_flat_supported(ctx->reads.max_len, ctx->reads.stride, ctx->K);
    if (flat) { int rc = ensure_flat_buffers(ctx, q_hi - q_lo); if (rc) return rc; }
    unsigned long long st_before[ST_COUNT] = {};
    if (flat) { // (several ranges may be searched one after the other: keep their counters)
        CK(cudaMemcpyAsync(st_before, ctx->d_stats_c, sizeof st_before, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    int rc = record(ctx, EV_CONT_K0);
    if (rc) return rc;
    for (int attempt = 0;; attempt++) {
        CK(cudaMemsetAsync(ctx->d_cursors + CUR_WORK, 0, 3 * sizeof(unsigned long long), ctx->stream));
        if (flat) {
            CK(cudaMemsetAsync(ctx->d_cursors + CUR_CANDS, 0, sizeof(unsigned long long), ctx->stream));
            if (attempt) CK(cudaMemcpyAsync(ctx->d_stats_c, st_before, sizeof st_before, cudaMemcpyHostToDevice, ctx->stream));
            p.cands = ctx->d_cands; p.cands_cap = ctx->cands_cap; p.cands_cursor = ctx->d_cursors + CUR_CANDS; p.batchinfo = ctx->d_batchinfo;
        }
        if (q_hi > q_lo) CK(launch_search_contained(p, ctx->num_sms, ctx->stream));
        if (!flat) break;
        unsigned long long cur[CUR_COUNT] = {}, st[ST_COUNT];
        CK(cudaMemcpyAsync(cur, ctx->d_cursors, sizeof cur, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(st, ctx->d_stats_c, sizeof st, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        if (!st[ST_OVERFLOW]) break;
        if (attempt >= 2) return fail(ctx, DISCO_E_NOMEM, "containment pass: candidate buffer overflow after retries (%llu entries needed)", cur[CUR_CANDS]);
        const uint64_t need = cur[CUR_CANDS] + edges_flat_slack(ctx->num_sms);
        dfree(ctx->d_cands);
        ctx->cands_cap = 0;
        CK(cudaMalloc(&ctx->d_cands, need * sizeof(uint64_t)));
        ctx->cands_cap = need;
        // (keys written by the incomplete attempt are valid minima: they can only be confirmed by the repeat)
    }
    if (flat) CK(cudaMemsetAsync(ctx->d_rowinfo + q_lo, 0, (q_hi - q_lo) * sizeof(uint64_t), ctx->stream)); // fall-back flags
    if ((rc = record(ctx, EV_CONT_K1))) return rc;
    return record(ctx, EV_CONTAINED);
}

int disco_gpu_phase_finish_contained(disco_ctx *ctx)
{
    if (!ctx || !ctx->begun) return fail(ctx, DISCO_E_ARG, "call disco_gpu_begin first");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemsetAsync(ctx->d_cursors + CUR_NCONTAINED, 0, 2 * sizeof(unsigned long long), ctx->stream)); // + CUR_CROWS
    CK(launch_contained_finish(ctx->d_best, ctx->reads.n, ctx->d_bits, ctx->d_cursors + CUR_NCONTAINED, ctx->stream));
    unsigned long long nc = 0, full = 0;
    CK(cudaMemcpyAsync(&nc, ctx->d_cursors + CUR_NCONTAINED, sizeof nc, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(&full, ctx->d_cursors + CUR_TABLE_FULL, sizeof full, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (full) return fail(ctx, DISCO_E_LIMIT, "hash table%s full: %llu buckets cannot hold this rank's keys (skewed k-mers)", ctx->tw() > 1 ? " shard" : "", (unsigned long long)ctx->nbuckets);
    ctx->n_contained = nc;
    if (nc > ctx->crows_cap) {
        dfree(ctx->d_crows);
        ctx->crows_cap = 0;
        CK(cudaMalloc(&ctx->d_crows, nc * sizeof(disco_crow)));
        ctx->crows_cap = nc;
    }
    if (nc) {
        CK(launch_contained_rows(ctx->d_best, ctx->reads, ctx->K, ctx->d_crows, ctx->d_cursors + CUR_CROWS, ctx->stream));
    }
    ctx->have_contained = true;
    ctx->stats.n_contained = nc;
    return record(ctx, EV_FINISH);
}

// Edge pass over the query reads [part_lo, part_hi) of this context's share [q_lo, q_hi).  part_lo == q_lo starts the
// pass (adjacency cursor and counters reset, buffer sized for the whole share); later parts append.  A caller that
// overlaps communication with compute (multigpu.py) runs the share in a few parts; everybody else runs one.
int disco_gpu_phase_edges_part(disco_ctx *ctx, uint64_t q_lo, uint64_t q_hi, uint64_t part_lo, uint64_t part_hi)
{
    if (!ctx || !ctx->begun || !ctx->have_contained) return fail(ctx, DISCO_E_ARG, "containment pass not finished");
    if (q_lo > q_hi || q_hi > ctx->reads.n || part_lo < q_lo || part_hi > q_hi || part_lo > part_hi) return fail(ctx, DISCO_E_ARG, "bad query range");
    if (ctx->tw() > 1 && !ctx->peer_table.ready) return fail(ctx, DISCO_E_ARG, "sharded table: import the peers' shards first (disco_gpu_import_peers)");
    CK(cudaSetDevice(ctx->device));
    const bool first = part_lo == q_lo;
    if (!first && !ctx->have_edges) return fail(ctx, DISCO_E_ARG, "edge pass parts must start at q_lo");
    const uint64_t nq = q_hi - q_lo, np = part_hi - part_lo;
    SearchParams p{};
    p.reads = ctx->reads; p.table = table_view(ctx);
    p.K = ctx->K; p.cap = ctx->cap; p.q_lo = part_lo; p.q_hi = part_hi;
    p.work_counter = ctx->d_cursors + CUR_WORK; p.stats = ctx->d_stats_e;
    p.contained_bits = ctx->d_bits; p.rows_cursor = ctx->d_cursors + CUR_ROWS; p.rowinfo = ctx->d_rowinfo;
    p.skip_contained = ctx->table_has_contained ? 1 : 0;
    // adjacency capacity: start from 48 entries per query read (30x, 150 bp, minOverlap 50 needs ~33), bounded by free
    // memory; the kernel keeps counting on overflow so that one retry with the exact size always succeeds
    if (!ctx->d_rows) {
        size_t free_b = 0, total_b = 0;
        CK(cudaMemGetInfo(&free_b, &total_b));
        uint64_t want = std::max<uint64_t>(nq * entries_per_read_guess(), getenv("DISCO_ENTRIES_PER_READ") ? 1 << 10 : 1 << 20);
        const uint64_t lim = (uint64_t)(free_b * 0.6) / sizeof(uint64_t);
        if (want > lim) want = std::max<uint64_t>(lim, 1 << 16);
        CK(cudaMalloc(&ctx->d_rows, want * sizeof(uint64_t)));
        ctx->rows_cap = want;
    }
    ctx->d_rows_active = ctx->d_rows;
    // flat kernels (short reads): candidate lists of this launch + four segment words per 32-read batch
    const bool flat = edges_flat_supported(ctx->reads.max_len, ctx->reads.stride, ctx->K);
    if (flat) { int rc = ensure_flat_buffers(ctx, np); if (rc) return rc; }
    const uint64_t cursor_before = first ? 0 : ctx->rows_used;
    unsigned long long st_before[ST_COUNT] = {};
    if (!first) {
        CK(cudaMemcpyAsync(st_before, ctx->d_stats_e, sizeof st_before, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    for (int attempt = 0;; attempt++) {
        CK(cudaMemsetAsync(ctx->d_cursors + CUR_WORK, 0, 3 * sizeof(unsigned long long), ctx->stream)); // 3 work counters
        CK(cudaMemsetAsync(ctx->d_cursors + CUR_CANDS, 0, sizeof(unsigned long long), ctx->stream));
        const unsigned long long cb = cursor_before;
        CK(cudaMemcpyAsync(ctx->d_cursors + CUR_ROWS, &cb, sizeof cb, cudaMemcpyHostToDevice, ctx->stream));
        if (first) CK(cudaMemsetAsync(ctx->d_stats_e, 0, ST_COUNT * sizeof(unsigned long long), ctx->stream));
        else if (attempt) CK(cudaMemcpyAsync(ctx->d_stats_e, st_before, sizeof st_before, cudaMemcpyHostToDevice, ctx->stream));
        if (attempt) CK(cudaMemsetAsync(ctx->d_rowinfo + part_lo, 0, np * sizeof(uint64_t), ctx->stream));
        p.rows = ctx->d_rows; p.rows_cap = ctx->rows_cap;
        p.cands = flat ? ctx->d_cands : nullptr; p.cands_cap = ctx->cands_cap;
        p.cands_cursor = ctx->d_cursors + CUR_CANDS; p.batchinfo = ctx->d_batchinfo;
        { int rc = record(ctx, EV_EDGES_K0); if (rc) return rc; }
        if (np) {
            CK(launch_search_edges(p, ctx->num_sms, ctx->stream, ctx->ev[EV_PROBE_K1], ctx->ev[EV_VERIFY_K1]));
            ctx->ev_done[EV_PROBE_K1] = ctx->ev_done[EV_VERIFY_K1] = !getenv("DISCO_FUSED");
        }
        { int rc = record(ctx, EV_EDGES_K1); if (rc) return rc; }
        unsigned long long cur[CUR_COUNT] = {}, st[ST_COUNT];
        CK(cudaMemcpyAsync(cur, ctx->d_cursors, sizeof cur, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(st, ctx->d_stats_e, sizeof st, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        if (cur[CUR_TABLE_FULL]) return fail(ctx, DISCO_E_LIMIT, "hash table%s full: %llu buckets cannot hold this rank's keys (skewed k-mers)", ctx->tw() > 1 ? " shard" : "", (unsigned long long)ctx->nbuckets);
        ctx->stats.raw_directed_edges = st[ST_ENTRIES];
        ctx->stats.max_degree = st[ST_MAXDEG];
        if (!st[ST_OVERFLOW]) {
            ctx->rows_used = cur[CUR_ROWS];
            if (first) ctx->acc_probe = ctx->acc_verify = ctx->acc_exact = ctx->acc_edges = 0.f;
            float t = 0.f;
            if (cudaEventElapsedTime(&t, ctx->ev[EV_EDGES_K0], ctx->ev[EV_EDGES_K1]) == cudaSuccess) ctx->acc_edges += t;
            if (ctx->ev_done[EV_PROBE_K1] && np) {
                if (cudaEventElapsedTime(&t, ctx->ev[EV_EDGES_K0], ctx->ev[EV_PROBE_K1]) == cudaSuccess) ctx->acc_probe += t;
                if (cudaEventElapsedTime(&t, ctx->ev[EV_PROBE_K1], ctx->ev[EV_VERIFY_K1]) == cudaSuccess) ctx->acc_verify += t;
                if (cudaEventElapsedTime(&t, ctx->ev[EV_VERIFY_K1], ctx->ev[EV_EDGES_K1]) == cudaSuccess) ctx->acc_exact += t;
            }
            break;
        }
        if (attempt >= 3) return fail(ctx, DISCO_E_NOMEM, "edge pass buffers overflow after retries (%llu adjacency entries, %llu candidates needed)", cur[CUR_ROWS], cur[CUR_CANDS]);
        if (st[ST_OVERFLOW] & 2) { // candidate lists: nothing to keep, the cursor says what this part needs
            const uint64_t need = cur[CUR_CANDS] + edges_flat_slack(ctx->num_sms);
            dfree(ctx->d_cands);
            ctx->cands_cap = 0;
            CK(cudaMalloc(&ctx->d_cands, need * sizeof(uint64_t)));
            ctx->cands_cap = need;
        }
        if (st[ST_OVERFLOW] & 1) {
            if (!ctx->own_rows) return fail(ctx, DISCO_E_NOMEM, "adopted adjacency buffer too small: %llu entries needed, %llu given", cur[CUR_ROWS], (unsigned long long)ctx->rows_cap);
            // grow (keeping earlier parts) and redo this part.  Slices are handed out per warp, so the slack differs between
            // runs: add one slice per resident warp; scale for the parts still to come
            uint64_t need = cur[CUR_ROWS] + (uint64_t)ctx->num_sms * 64 * 1024;
            if (np && np < nq) need += (cur[CUR_ROWS] - cursor_before) * ((q_hi - part_hi) / np + 1);
            uint64_t *nr = nullptr;
            if (!cursor_before) dfree(ctx->d_rows); // nothing to keep: do not hold both buffers at once
            CK(cudaMalloc(&nr, need * sizeof(uint64_t)));
            if (cursor_before) CK(cudaMemcpyAsync(nr, ctx->d_rows, cursor_before * sizeof(uint64_t), cudaMemcpyDeviceToDevice, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            dfree(ctx->d_rows);
            ctx->d_rows = ctx->d_rows_active = nr;
            ctx->rows_cap = need;
        }
        if (!first) st_before[ST_OVERFLOW] = 0;
    }
    ctx->stats.edge_capacity = ctx->rows_cap;
    ctx->have_edges = true;
    return record(ctx, EV_EDGES);
}

int disco_gpu_phase_edges(disco_ctx *ctx, uint64_t q_lo, uint64_t q_hi) { return disco_gpu_phase_edges_part(ctx, q_lo, q_hi, q_lo, q_hi); }

namespace {
int reduce_params(disco_ctx *ctx, uint64_t u_lo, uint64_t u_hi, ReduceParams &p)
{
    if (!ctx || !ctx->have_edges) return fail(ctx, DISCO_E_ARG, "edge pass not finished");
    if (u_lo > u_hi || u_hi > ctx->reads.n) return fail(ctx, DISCO_E_ARG, "bad node range");
    if (ctx->shard_world > 1 && !ctx->peer_rows.ready) return fail(ctx, DISCO_E_ARG, "sharded adjacency: import the peers' rows first (disco_gpu_import_peers)");
    p = ReduceParams{};
    p.reads = ctx->reads; p.rows = ctx->d_rows_active; p.rowinfo = ctx->d_rowinfo; p.u_lo = u_lo; p.u_hi = u_hi;
    p.work_counter = ctx->d_cursors + CUR_WORK; p.stats = ctx->d_stats_e;
    p.peer_rows = ctx->peer_rows.d_ptrs; p.bounds = ctx->d_bounds; p.world = ctx->shard_world;
    if (const char *e = getenv("DISCO_PEER_LOAD")) p.peer_load = atoi(e);
    p.maxdeg = (int)std::max<uint64_t>(ctx->stats.max_degree, 1);
    if ((size_t)p.maxdeg * 25 + 512 > 200 * 1024) return fail(ctx, DISCO_E_LIMIT, "max degree %d too large for the reduction kernel", p.maxdeg);
    return DISCO_OK;
}
} // namespace

// markTransitiveEdges for the nodes [u_lo, u_hi): sets the eliminated bit of their own entries
int disco_gpu_phase_reduce_mark(disco_ctx *ctx, uint64_t u_lo, uint64_t u_hi)
{
    ReduceParams p;
    int rc = reduce_params(ctx, u_lo, u_hi, p);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemsetAsync(ctx->d_cursors + CUR_WORK, 0, sizeof(unsigned long long), ctx->stream));
    if ((rc = record(ctx, EV_MARK_K0))) return rc;
    if (u_hi > u_lo && ctx->stats.raw_directed_edges) CK(launch_reduce_mark(p, ctx->num_sms, ctx->stream));
    return record(ctx, EV_MARK);
}

// removeTransitiveEdges + canonical selection for the nodes [u_lo, u_hi); every node's marks must be in place (in the
// sharded mode: on every GPU -- the caller puts a barrier between the two calls)
int disco_gpu_phase_reduce_emit(disco_ctx *ctx, uint64_t u_lo, uint64_t u_hi)
{
    ReduceParams p;
    int rc = reduce_params(ctx, u_lo, u_hi, p);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    if (!ctx->d_edges) {
        const uint64_t want = std::max<uint64_t>(ctx->stats.raw_directed_edges / 4, 1 << 16);
        CK(cudaMalloc(&ctx->d_edges, want * sizeof(disco_edge)));
        ctx->edges_cap = want;
    }
    for (int attempt = 0;; attempt++) {
        CK(cudaMemsetAsync(ctx->d_cursors + CUR_WORK, 0, sizeof(unsigned long long), ctx->stream));
        CK(cudaMemsetAsync(ctx->d_cursors + CUR_EDGES, 0, sizeof(unsigned long long), ctx->stream));
        // the emission kernel's own counters: zeroed per attempt, so that a retry after an edge-buffer overflow counts once
        CK(cudaMemsetAsync(ctx->d_stats_e + ST_MULTI_OVERLAP, 0, (ST_COUNT - ST_MULTI_OVERLAP) * sizeof(unsigned long long), ctx->stream));
        p.edges_out = ctx->d_edges; p.edges_cap = ctx->edges_cap; p.edges_cursor = ctx->d_cursors + CUR_EDGES;
        p.edges_out2 = ctx->sink_dev; p.edges_cap2 = ctx->sink_cap;
        ctx->sink_filled = false;
        if ((rc = record(ctx, EV_EMIT_K0))) return rc;
        if (u_hi > u_lo && ctx->stats.raw_directed_edges) CK(launch_reduce_emit(p, ctx->num_sms, ctx->stream));
        if ((rc = record(ctx, EV_EMIT))) return rc;
        unsigned long long ne = 0;
        CK(cudaMemcpyAsync(&ne, ctx->d_cursors + CUR_EDGES, sizeof ne, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->n_edges = ne;
        if (ne <= ctx->edges_cap) break;
        if (attempt >= 1) return fail(ctx, DISCO_E_NOMEM, "edge buffer overflow after retry");
        dfree(ctx->d_edges);
        CK(cudaMalloc(&ctx->d_edges, ne * sizeof(disco_edge)));
        ctx->edges_cap = ne;
    }
    ctx->sink_filled = ctx->sink_dev != nullptr && ctx->n_edges <= ctx->sink_cap;
    ctx->stats.n_edges = ctx->n_edges;
    ctx->have_reduced = true;
    return DISCO_OK;
}

int disco_gpu_phase_reduce(disco_ctx *ctx, uint64_t u_lo, uint64_t u_hi)
{
    const int rc = disco_gpu_phase_reduce_mark(ctx, u_lo, u_hi);
    return rc ? rc : disco_gpu_phase_reduce_emit(ctx, u_lo, u_hi);
}

// ---- key-sharded mode (Mode B) --------------------------------------------------------------------------------------
int disco_gpu_set_shard(disco_ctx *ctx, uint32_t world, uint32_t rank) { return disco_gpu_set_partition(ctx, world, rank, 1); }

int disco_gpu_set_partition(disco_ctx *ctx, uint32_t world, uint32_t rank, int shard_table)
{
    if (!ctx) return DISCO_E_ARG;
    if (world < 1 || world > DISCO_MAX_SHARDS || rank >= world) return fail(ctx, DISCO_E_ARG, "bad shard %u of %u (at most %d)", rank, world, DISCO_MAX_SHARDS);
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    if (world != ctx->shard_world || rank != ctx->shard_rank || (shard_table != 0) != ctx->table_sharded) {
        close_peers(ctx->peer_table); close_peers(ctx->peer_rows);
        free_run_buffers(ctx); // the table is sized per shard
    }
    ctx->shard_world = world; ctx->shard_rank = rank; ctx->table_sharded = shard_table != 0;
    if (world > 1) {
        if (!ctx->peer_table.d_ptrs) CK(cudaMalloc(&ctx->peer_table.d_ptrs, DISCO_MAX_SHARDS * sizeof(uint64_t *)));
        if (!ctx->peer_rows.d_ptrs) CK(cudaMalloc(&ctx->peer_rows.d_ptrs, DISCO_MAX_SHARDS * sizeof(uint64_t *)));
        if (!ctx->peer_keys.d_ptrs) CK(cudaMalloc(&ctx->peer_keys.d_ptrs, DISCO_MAX_SHARDS * sizeof(uint64_t *)));
        if (!ctx->d_bounds) CK(cudaMalloc(&ctx->d_bounds, (DISCO_MAX_SHARDS + 1) * sizeof(uint64_t)));
    }
    return DISCO_OK;
}

int disco_gpu_export_mem(disco_ctx *ctx, int which, void *handle_out)
{
    if (!ctx || !handle_out) return DISCO_E_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == DISCO_IPC_HANDLE_BYTES, "IPC handle size");
    void *ptr = which == DISCO_MEM_TABLE ? (void *)ctx->d_slots : which == DISCO_MEM_ROWS ? (void *)ctx->d_rows : nullptr;
    if (!ptr) return fail(ctx, DISCO_E_ARG, "nothing to export (which = %d): buffer not allocated yet", which);
    CK(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, ptr));
    memcpy(handle_out, &h, sizeof h);
    return DISCO_OK;
}

namespace {
int check_import(disco_ctx *ctx, int which, const void *src, const uint64_t *bounds)
{
    if (!ctx || !src) return DISCO_E_ARG;
    if ((which == DISCO_MEM_TABLE ? ctx->tw() : ctx->shard_world) < 2) return fail(ctx, DISCO_E_ARG, "not in sharded mode (disco_gpu_set_shard / disco_gpu_set_partition)");
    if (which != DISCO_MEM_TABLE && which != DISCO_MEM_ROWS) return fail(ctx, DISCO_E_ARG, "bad buffer selector %d", which);
    if (which == DISCO_MEM_ROWS && !bounds) return fail(ctx, DISCO_E_ARG, "the adjacency needs the ranks' read-id bounds");
    if (!(which == DISCO_MEM_TABLE ? (void *)ctx->d_slots : (void *)ctx->d_rows)) return fail(ctx, DISCO_E_ARG, "own buffer not allocated yet");
    if (which == DISCO_MEM_ROWS)
        for (uint32_t r = 0; r < ctx->shard_world; r++)
            if (bounds[r] > bounds[r + 1] || bounds[r + 1] > ctx->reads.n) return fail(ctx, DISCO_E_ARG, "bad read-id bounds");
    return DISCO_OK;
}

int publish_peers(disco_ctx *ctx, int which, const uint64_t *const *ptrs, const uint64_t *bounds)
{
    disco_ctx::PeerSet &ps = which == DISCO_MEM_TABLE ? ctx->peer_table : ctx->peer_rows;
    CK(cudaMemcpy(ps.d_ptrs, ptrs, DISCO_MAX_SHARDS * sizeof(uint64_t *), cudaMemcpyHostToDevice));
    if (which == DISCO_MEM_ROWS) CK(cudaMemcpy(ctx->d_bounds, bounds, (ctx->shard_world + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice));
    ps.ready = true;
    return DISCO_OK;
}
} // namespace

int disco_gpu_import_peers(disco_ctx *ctx, int which, const void *handles, const uint64_t *bounds)
{
    int rc = check_import(ctx, which, handles, bounds);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    disco_ctx::PeerSet &ps = which == DISCO_MEM_TABLE ? ctx->peer_table : ctx->peer_rows;
    const cudaIpcMemHandle_t *hs = static_cast<const cudaIpcMemHandle_t *>(handles);
    const uint64_t *ptrs[DISCO_MAX_SHARDS] = {};
    for (uint32_t r = 0; r < ctx->shard_world; r++) {
        if (r == ctx->shard_rank) { ptrs[r] = which == DISCO_MEM_TABLE ? ctx->d_slots : ctx->d_rows; continue; }
        if (ps.opened[r] && memcmp(&ps.handle[r], &hs[r], sizeof(cudaIpcMemHandle_t)) != 0) { // the peer reallocated
            cudaIpcCloseMemHandle(ps.opened[r]);
            ps.opened[r] = nullptr;
        }
        if (!ps.opened[r]) {
            CK(cudaIpcOpenMemHandle(&ps.opened[r], hs[r], cudaIpcMemLazyEnablePeerAccess));
            ps.handle[r] = hs[r];
        }
        ptrs[r] = static_cast<const uint64_t *>(ps.opened[r]);
    }
    return publish_peers(ctx, which, ptrs, bounds);
}

// the same for shards that live in this process (one host process driving several contexts): plain device pointers; the
// caller has enabled peer access between the devices involved
int disco_gpu_import_peer_ptrs(disco_ctx *ctx, int which, const void *const *device_ptrs, const uint64_t *bounds)
{
    int rc = check_import(ctx, which, device_ptrs, bounds);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    close_peers(which == DISCO_MEM_TABLE ? ctx->peer_table : ctx->peer_rows);
    const uint64_t *ptrs[DISCO_MAX_SHARDS] = {};
    for (uint32_t r = 0; r < ctx->shard_world; r++) {
        ptrs[r] = r == ctx->shard_rank ? (which == DISCO_MEM_TABLE ? ctx->d_slots : ctx->d_rows) : static_cast<const uint64_t *>(device_ptrs[r]);
        if (!ptrs[r]) return fail(ctx, DISCO_E_ARG, "NULL pointer for shard %u", r);
    }
    return publish_peers(ctx, which, ptrs, bounds);
}

void *disco_gpu_dev_table(disco_ctx *ctx) { return ctx ? ctx->d_slots : nullptr; }
uint64_t disco_gpu_table_words(disco_ctx *ctx) { return ctx && ctx->begun ? ctx->nbuckets * 4 : 0; }

// Caller-owned device memory for the table shard (after disco_gpu_begin, at least disco_gpu_table_words u64) or the
// adjacency (before the edge pass; n_u64 = capacity in entries, never grown here: the edge pass fails with
// DISCO_E_NOMEM when it is too small).  For memory that other GPUs map by other means than CUDA IPC handles (symmetric
// memory: VMM allocations with 2 MB pages).
int disco_gpu_adopt_buffer(disco_ctx *ctx, int which, void *d_ptr, uint64_t n_u64)
{
    if (!ctx || !d_ptr) return DISCO_E_ARG;
    if (!ctx->begun) return fail(ctx, DISCO_E_ARG, "call disco_gpu_begin first");
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    if (which == DISCO_MEM_TABLE) {
        if (n_u64 < ctx->nbuckets * 4) return fail(ctx, DISCO_E_ARG, "table buffer too small: %llu u64 needed", (unsigned long long)(ctx->nbuckets * 4));
        if (ctx->own_slots) dfree(ctx->d_slots);
        ctx->d_slots = static_cast<uint64_t *>(d_ptr);
        ctx->own_slots = false;
        ctx->peer_table.ready = false;
    } else if (which == DISCO_MEM_ROWS) {
        if (n_u64 < (1u << 16)) return fail(ctx, DISCO_E_ARG, "adjacency buffer too small");
        if (ctx->own_rows) dfree(ctx->d_rows);
        ctx->d_rows = ctx->d_rows_active = static_cast<uint64_t *>(d_ptr);
        ctx->rows_cap = n_u64;
        ctx->rows_used = 0;
        ctx->own_rows = false;
        ctx->have_edges = ctx->have_reduced = false;
        ctx->peer_rows.ready = false;
    } else {
        return fail(ctx, DISCO_E_ARG, "bad buffer selector %d", which);
    }
    return DISCO_OK;
}

// ---- several GPUs driven by ONE process (what `buildG -g 0,1,...` uses): Mode B with one host thread per context.
// The ranks' buffers are mapped by peer access (same address space), the containment keys are min-reduced by a kernel
// that reads the peers' arrays, the row infos are copied range by range -- no NCCL, no MPI.
namespace {
struct HostBarrier {
    std::mutex m;
    std::condition_variable cv;
    unsigned n, waiting = 0, generation = 0;
    explicit HostBarrier(unsigned n_) : n(n_) {}
    void wait()
    {
        std::unique_lock<std::mutex> lk(m);
        const unsigned g = generation;
        if (++waiting == n) { waiting = 0; generation++; cv.notify_all(); }
        else cv.wait(lk, [&] { return generation != g; });
    }
};
} // namespace

int disco_gpu_build_graph_multi(disco_ctx *const *ctxs, uint32_t world, uint32_t min_overlap, uint32_t max_edge_per_kmer)
{
    if (!ctxs || world < 1 || world > DISCO_MAX_SHARDS) return DISCO_E_ARG;
    for (uint32_t r = 0; r < world; r++) if (!ctxs[r]) return DISCO_E_ARG;
    if (world == 1) {
        int rc = disco_gpu_set_shard(ctxs[0], 1, 0);
        return rc ? rc : disco_gpu_build_graph(ctxs[0], min_overlap, max_edge_per_kmer);
    }
    disco_ctx *c0 = ctxs[0];
    const uint64_t n = c0->reads.n;
    for (uint32_t r = 0; r < world; r++) {
        disco_ctx *c = ctxs[r];
        if (!c->d_words) return fail(c, DISCO_E_ARG, "load reads first");
        if (c->reads.n != n || c->reads.max_len != c0->reads.max_len || c->reads.min_len != c0->reads.min_len)
            return fail(c, DISCO_E_ARG, "every context must hold the same read set (reads are replicated)");
        for (uint32_t q = 0; q < r; q++) if (ctxs[q] == c) return fail(c, DISCO_E_ARG, "the same context twice");
    }
    // peer access between every pair of distinct devices
    for (uint32_t a = 0; a < world; a++)
        for (uint32_t b = 0; b < world; b++) {
            const int da = ctxs[a]->device, db = ctxs[b]->device;
            if (da == db) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, da, db);
            if (!can) return fail(ctxs[a], DISCO_E_CUDA, "device %d cannot access device %d: no peer path", da, db);
            cudaSetDevice(da);
            const cudaError_t e = cudaDeviceEnablePeerAccess(db, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(ctxs[a], DISCO_E_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d): %s", da, db, cudaGetErrorString(e));
            cudaGetLastError();
        }
    uint64_t bounds[DISCO_MAX_SHARDS + 1];
    for (uint32_t r = 0; r <= world; r++) bounds[r] = ((uint64_t)r * n) / world; // n < 2^31: no overflow; = multigpu.partition
    const void *tables[DISCO_MAX_SHARDS] = {}, *rows[DISCO_MAX_SHARDS] = {}, *keys[DISCO_MAX_SHARDS] = {};
    uint64_t maxdeg[DISCO_MAX_SHARDS] = {};
    int rcs[DISCO_MAX_SHARDS] = {};
    std::atomic<bool> failed{false};
    HostBarrier bar(world);

    auto body = [&](uint32_t r) {
        disco_ctx *ctx = ctxs[r];
        const uint64_t lo = bounds[r], hi = bounds[r + 1];
        auto run = [&](auto &&fn) { // skipped once any rank failed; the barriers are always kept
            if (failed.load()) return;
            const int rc = fn();
            if (rc) { rcs[r] = rc; failed.store(true); }
        };
        auto cuda = [&](cudaError_t e, const char *what) -> int {
            return e == cudaSuccess ? DISCO_OK : fail(ctx, DISCO_E_CUDA, "%s: %s", what, cudaGetErrorString(e));
        };
        run([&] { return disco_gpu_set_shard(ctx, world, r); });
        run([&] { return disco_gpu_begin(ctx, min_overlap, max_edge_per_kmer); });
        run([&] { return disco_gpu_phase_table(ctx, 0); });
        run([&] { return disco_gpu_sync(ctx); });
        tables[r] = ctx->d_slots; keys[r] = ctx->d_best;
        bar.wait();                                   // every shard complete, every pointer published
        run([&] { return disco_gpu_import_peer_ptrs(ctx, DISCO_MEM_TABLE, tables, nullptr); });
        run([&] { return cuda(cudaMemcpy(ctx->peer_keys.d_ptrs, keys, sizeof keys, cudaMemcpyHostToDevice), "publish key pointers"); });
        run([&] { return disco_gpu_phase_contained(ctx, lo, hi); });
        run([&] { return disco_gpu_sync(ctx); });
        bar.wait();                                   // every rank's keys written
        run([&] { return cuda(launch_min_keys(ctx->d_best, ctx->peer_keys.d_ptrs, world, r, n, ctx->stream), "min-reduce of the containment keys"); });
        run([&] { return disco_gpu_phase_finish_contained(ctx); });
        bar.wait();                                   // nobody probes the old table (or reads our keys) any more
        run([&] { return disco_gpu_phase_table(ctx, 1); });
        run([&] { return disco_gpu_sync(ctx); });
        bar.wait();
        run([&] { return disco_gpu_phase_edges(ctx, lo, hi); });
        rows[r] = ctx->d_rows; maxdeg[r] = ctx->stats.max_degree;
        bar.wait();                                   // every rank's rows and row infos complete
        run([&] {
            for (uint32_t q = 0; q < world; q++) {    // row infos: each rank's own range is final, copy it over
                if (q == r || bounds[q + 1] == bounds[q]) continue;
                const int rc = cuda(cudaMemcpyPeerAsync(ctx->d_rowinfo + bounds[q], ctx->device, ctxs[q]->d_rowinfo + bounds[q], ctxs[q]->device,
                                                        (bounds[q + 1] - bounds[q]) * sizeof(uint64_t), ctx->stream), "row info exchange");
                if (rc) return rc;
            }
            return DISCO_OK;
        });
        run([&] { return disco_gpu_set_max_degree(ctx, *std::max_element(maxdeg, maxdeg + world)); });
        run([&] { return disco_gpu_import_peer_ptrs(ctx, DISCO_MEM_ROWS, rows, bounds); });
        run([&] { return disco_gpu_phase_reduce_mark(ctx, lo, hi); });
        run([&] { return disco_gpu_sync(ctx); });
        bar.wait();                                   // emission reads the marks of remote neighbours
        run([&] { return disco_gpu_phase_reduce_emit(ctx, lo, hi); });
        bar.wait();                                   // nobody reads a peer's rows after this point
    };
    std::vector<std::thread> threads;
    for (uint32_t r = 1; r < world; r++) threads.emplace_back(body, r);
    body(0);
    for (auto &t : threads) t.join();
    for (uint32_t r = 0; r < world; r++) if (rcs[r]) return rcs[r];
    return DISCO_OK;
}

int disco_gpu_build_graph(disco_ctx *ctx, uint32_t min_overlap, uint32_t max_edge_per_kmer)
{
    int rc;
    if ((rc = disco_gpu_begin(ctx, min_overlap, max_edge_per_kmer))) return rc;
    const uint64_t n = ctx->reads.n;
    if ((rc = disco_gpu_phase_table(ctx, 0))) return rc;
    if ((rc = disco_gpu_phase_contained(ctx, 0, n))) return rc;
    if ((rc = disco_gpu_phase_finish_contained(ctx))) return rc;
    // Second table without the contained reads: 1.4 ms per 10M reads buys a 2.6 ms faster probe kernel (no bitmap check
    // per tag match, shorter chains).  DISCO_SINGLE_TABLE=1 skips it (the edge pass then drops contained candidates
    // through the bitmap) -- the better trade when the table is replicated on many GPUs, see multigpu.py.
    if (!getenv("DISCO_SINGLE_TABLE") && (rc = disco_gpu_phase_table(ctx, 1))) return rc;
    if ((rc = disco_gpu_phase_edges(ctx, 0, n))) return rc;
    if ((rc = disco_gpu_phase_reduce(ctx, 0, n))) return rc;
    return DISCO_OK;
}

// ---- results ------------------------------------------------------------------------------------------------------
int disco_gpu_counts(disco_ctx *ctx, uint64_t *n_contained, uint64_t *n_edges)
{
    if (!ctx) return DISCO_E_ARG;
    if (n_contained) *n_contained = ctx->n_contained;
    if (n_edges) *n_edges = ctx->n_edges;
    return DISCO_OK;
}

int disco_gpu_get_contained(disco_ctx *ctx, disco_crow *rows, uint64_t capacity, uint64_t *n_written)
{
    if (!ctx || !ctx->have_contained) return fail(ctx, DISCO_E_ARG, "containment pass not finished");
    if (capacity < ctx->n_contained) return fail(ctx, DISCO_E_ARG, "capacity %llu < %llu rows", (unsigned long long)capacity, (unsigned long long)ctx->n_contained);
    CK(cudaSetDevice(ctx->device));
    const uint64_t n = ctx->n_contained;
    if (n) {
        CK(cudaMemcpyAsync(rows, ctx->d_crows, n * sizeof(disco_crow), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    if (n_written) *n_written = n;
    return DISCO_OK;
}

// Only the rows whose contained read lies in [read_lo, read_hi): every context of a multi-GPU run holds all rows; each
// rank hands out (writes) those of its own read range, as it does with the edges.  Order not defined.
int disco_gpu_get_contained_range(disco_ctx *ctx, uint64_t read_lo, uint64_t read_hi, disco_crow *rows, uint64_t capacity, uint64_t *n_written)
{
    if (!ctx || !ctx->have_contained) return fail(ctx, DISCO_E_ARG, "containment pass not finished");
    if (read_lo > read_hi) return fail(ctx, DISCO_E_ARG, "bad read range");
    CK(cudaSetDevice(ctx->device));
    const uint64_t n = ctx->n_contained;
    unsigned long long k = 0;
    if (n) {
        if (ctx->crows_part_cap < n) {
            dfree(ctx->d_crows_part);
            ctx->crows_part_cap = 0;
            CK(cudaMalloc(&ctx->d_crows_part, n * sizeof(disco_crow)));
            ctx->crows_part_cap = n;
        }
        CK(cudaMemsetAsync(ctx->d_cursors + CUR_CROWS, 0, sizeof(unsigned long long), ctx->stream));
        CK(launch_crows_in_range(ctx->d_crows, n, read_lo, read_hi, ctx->d_crows_part, ctx->d_cursors + CUR_CROWS, ctx->stream));
        CK(cudaMemcpyAsync(&k, ctx->d_cursors + CUR_CROWS, sizeof k, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        if (k > capacity) return fail(ctx, DISCO_E_ARG, "capacity %llu < %llu rows", (unsigned long long)capacity, k);
        if (k) {
            CK(cudaMemcpyAsync(rows, ctx->d_crows_part, k * sizeof(disco_crow), cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
        }
    }
    if (n_written) *n_written = k;
    return DISCO_OK;
}

int disco_gpu_get_edges(disco_ctx *ctx, disco_edge *edges, uint64_t capacity, uint64_t *n_written)
{
    if (!ctx || !ctx->have_reduced) return fail(ctx, DISCO_E_ARG, "reduction not finished");
    if (capacity < ctx->n_edges) return fail(ctx, DISCO_E_ARG, "capacity too small");
    CK(cudaSetDevice(ctx->device));
    if (ctx->n_edges && !(edges == ctx->sink_host && ctx->sink_filled)) { // (the sink already holds them: the kernel wrote it)
        CK(cudaMemcpyAsync(edges, ctx->d_edges, ctx->n_edges * sizeof(disco_edge), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    if (n_written) *n_written = ctx->n_edges;
    return DISCO_OK;
}

// The reduced edges put into (src, dst) order where they lie in HBM, so that disco_gpu_get_edges returns them sorted and
// the host has nothing left to sort (the emission kernel appends them in whatever order its warps finish).  The pinned
// sink, if any, keeps the emission order: after this call disco_gpu_get_edges copies from the device again.
int disco_gpu_sort_edges(disco_ctx *ctx)
{
    if (!ctx || !ctx->have_reduced) return fail(ctx, DISCO_E_ARG, "reduction not finished");
    CK(cudaSetDevice(ctx->device));
    unsigned long long launches = 0;
    const cudaError_t e = sort_edges_device(ctx->d_edges, ctx->n_edges, ctx->reads.n, ctx->stream, &launches);
    count_launches(launches);
    if (e != cudaSuccess) return fail(ctx, e == cudaErrorMemoryAllocation ? DISCO_E_NOMEM : DISCO_E_CUDA, "edge sort failed: %s", cudaGetErrorString(e));
    ctx->sink_filled = false;
    return DISCO_OK;
}

// ---- first consumer step on the device-resident edges (simplify.cu) ----------------------------------------------------
int disco_gpu_simplify(disco_ctx *ctx, uint32_t min_overlap, uint32_t min_reads, uint32_t min_len, uint64_t *n_edges, uint64_t *n_inner)
{
    if (!ctx || !ctx->have_reduced) return fail(ctx, DISCO_E_ARG, "reduction not finished");
    if (ctx->shard_world > 1) return fail(ctx, DISCO_E_ARG, "simplify works on one context's complete edge set (single-GPU runs)");
    CK(cudaSetDevice(ctx->device));
    dfree(ctx->d_cedges); dfree(ctx->d_inner);
    ctx->n_cedges = ctx->n_inner = 0;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0, ctx->stream));
    unsigned long long launches = 0;
    const cudaError_t e = run_simplify(ctx->d_edges, ctx->n_edges, ctx->reads.n, ctx->d_len, ctx->reads.uniform_len, min_overlap, min_reads, min_len,
                                       ctx->stream, &ctx->d_cedges, &ctx->n_cedges, &ctx->d_inner, &ctx->n_inner, &ctx->simp_rounds,
                                       &ctx->simp_removed, &ctx->simp_cycle, &launches);
    count_launches(launches);
    cudaEventRecord(e1, ctx->stream);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ctx->simp_ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (e != cudaSuccess) return fail(ctx, e == cudaErrorMemoryAllocation ? DISCO_E_NOMEM : DISCO_E_CUDA, "simplify failed: %s", cudaGetErrorString(e));
    if (n_edges) *n_edges = ctx->n_cedges;
    if (n_inner) *n_inner = ctx->n_inner;
    return DISCO_OK;
}

int disco_gpu_get_simplified(disco_ctx *ctx, disco_cedge *edges, uint64_t edge_capacity, uint64_t *inner, uint64_t inner_capacity)
{
    if (!ctx || !ctx->d_cedges) return fail(ctx, DISCO_E_ARG, "call disco_gpu_simplify first");
    if (edge_capacity < ctx->n_cedges || inner_capacity < ctx->n_inner) return fail(ctx, DISCO_E_ARG, "capacity too small");
    CK(cudaSetDevice(ctx->device));
    if (ctx->n_cedges) CK(cudaMemcpyAsync(edges, ctx->d_cedges, ctx->n_cedges * sizeof(disco_cedge), cudaMemcpyDeviceToHost, ctx->stream));
    if (ctx->n_inner) CK(cudaMemcpyAsync(inner, ctx->d_inner, ctx->n_inner * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return DISCO_OK;
}

int disco_gpu_simplify_stats(disco_ctx *ctx, uint64_t *rounds, uint64_t *removed_edges, uint64_t *cycle_edges, float *ms)
{
    if (!ctx) return DISCO_E_ARG;
    if (rounds) *rounds = ctx->simp_rounds;
    if (removed_edges) *removed_edges = ctx->simp_removed;
    if (cycle_edges) *cycle_edges = ctx->simp_cycle;
    if (ms) *ms = ctx->simp_ms;
    return DISCO_OK;
}

// The emission kernel writes every kept edge to this pinned host buffer as well (over PCIe, while it runs), so that the
// device-to-host copy of the result is off the critical path: disco_gpu_get_edges(ctx, host_pinned, ...) then returns at
// once.  NULL clears it.  A result larger than the capacity is simply not mirrored (get_edges copies as usual).
int disco_gpu_set_edge_sink(disco_ctx *ctx, disco_edge *host_pinned, uint64_t capacity)
{
    if (!ctx) return DISCO_E_ARG;
    CK(cudaSetDevice(ctx->device));
    ctx->sink_host = ctx->sink_dev = nullptr; ctx->sink_cap = 0; ctx->sink_filled = false;
    if (!host_pinned || !capacity) return DISCO_OK;
    void *dp = nullptr;
    if (cudaHostGetDevicePointer(&dp, host_pinned, 0) != cudaSuccess) { cudaGetLastError(); return fail(ctx, DISCO_E_ARG, "edge sink must be page-locked host memory (cudaHostAlloc / cudaHostRegister)"); }
    ctx->sink_host = host_pinned; ctx->sink_dev = static_cast<disco_edge *>(dp); ctx->sink_cap = capacity;
    return DISCO_OK;
}

int disco_gpu_get_row(disco_ctx *ctx, uint64_t read, disco_edge *out, uint64_t capacity, uint64_t *n_written)
{
    if (!ctx || !ctx->have_edges) return fail(ctx, DISCO_E_ARG, "edge pass not finished");
    if (read >= ctx->reads.n) return fail(ctx, DISCO_E_ARG, "read out of range");
    CK(cudaSetDevice(ctx->device));
    uint64_t ri = 0;
    CK(cudaMemcpyAsync(&ri, ctx->d_rowinfo + read, sizeof ri, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const uint32_t deg = rowinfo_deg(ri);
    if (deg > capacity) return fail(ctx, DISCO_E_ARG, "capacity too small");
    std::vector<uint64_t> e(deg);
    if (deg) {
        CK(cudaMemcpyAsync(e.data(), ctx->d_rows_active + rowinfo_start(ri), deg * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    for (auto &x : e) x &= ~kElimBit;
    std::sort(e.begin(), e.end()); // device rows are unsorted; (offset, neighbour, orientation) is the reference's list order
    for (uint32_t i = 0; i < deg; i++) {
        out[i].src = (uint32_t)read; out[i].dst = (uint32_t)entry_nbr(e[i]);
        out[i].offset = (uint32_t)entry_offset(e[i]); out[i].orient = (uint32_t)entry_orient(e[i]);
    }
    if (n_written) *n_written = deg;
    return DISCO_OK;
}

int disco_gpu_get_stats(disco_ctx *ctx, disco_stats *out)
{
    if (!ctx || !out) return DISCO_E_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    unsigned long long sc[ST_COUNT] = {}, se[ST_COUNT] = {};
    CK(cudaMemcpy(sc, ctx->d_stats_c, sizeof sc, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(se, ctx->d_stats_e, sizeof se, cudaMemcpyDeviceToHost));
    disco_stats &s = ctx->stats;
    s.probes_contained = sc[ST_PROBES]; s.buckets_contained = sc[ST_BUCKETS]; s.verified_contained = sc[ST_VERIFIED];
    s.probes_edges = se[ST_PROBES]; s.buckets_edges = se[ST_BUCKETS]; s.verified_edges = se[ST_VERIFIED];
    s.queries_contained = sc[ST_QUERIES]; s.queries_edges = se[ST_QUERIES];
    s.cap_fired = se[ST_CAP_FIRED]; s.slow_path_reads = se[ST_SLOW_READS];
    s.multi_overlap_pairs = se[ST_MULTI_OVERLAP]; s.one_sided_edges = se[ST_ONE_SIDED];
    s.reduce_rows_fetched = se[ST_ROWS_FETCHED] + se[ST_EMIT_ROWS]; s.reduce_entries_fetched = se[ST_ENTRIES_FETCHED] + se[ST_EMIT_ENTRIES];
    s.kernel_launches = launches_total() - ctx->launches_at_begin;
    s.mark_rows_fetched = se[ST_ROWS_FETCHED]; s.mark_entries_fetched = se[ST_ENTRIES_FETCHED];
    auto ms = [&](int a, int b) {
        float t = 0.f;
        if (ctx->ev_done[a] && ctx->ev_done[b]) cudaEventElapsedTime(&t, ctx->ev[a], ctx->ev[b]);
        return t;
    };
    s.ms_table_all = ms(EV_T0, EV_TABLE_ALL); s.ms_contained = ms(EV_TABLE_ALL, EV_CONTAINED);
    s.ms_finish_contained = ms(EV_CONTAINED, EV_FINISH); s.ms_table_nc = ms(EV_FINISH, EV_TABLE_NC);
    s.ms_edges = ctx->ev_done[EV_TABLE_NC] ? ms(EV_TABLE_NC, EV_EDGES) : ms(EV_FINISH, EV_EDGES); s.ms_mark = ms(EV_EDGES, EV_MARK); s.ms_emit = ms(EV_MARK, EV_EMIT);
    s.ms_total = ms(EV_T0, EV_EMIT);
    s.ms_edges_kernel = ctx->acc_edges; s.ms_contained_kernel = ms(EV_CONT_K0, EV_CONT_K1);
    s.ms_edges_probe = ctx->acc_probe; s.ms_edges_verify = ctx->acc_verify; s.ms_edges_exact = ctx->acc_exact;
    s.ms_mark_kernel = ms(EV_MARK_K0, EV_MARK); s.ms_emit_kernel = ms(EV_EMIT_K0, EV_EMIT);
    *out = s;
    return DISCO_OK;
}

// ---- multi-GPU plumbing ---------------------------------------------------------------------------------------------
void *disco_gpu_dev_contained_keys(disco_ctx *ctx) { return ctx ? ctx->d_best : nullptr; }
void *disco_gpu_dev_rowinfo(disco_ctx *ctx) { return ctx ? ctx->d_rowinfo : nullptr; }
void *disco_gpu_dev_rows(disco_ctx *ctx, uint64_t *n_entries)
{
    if (!ctx) return nullptr;
    if (n_entries) *n_entries = ctx->rows_used;
    return ctx->d_rows;
}

// Sparse form of the containment-key exchange (multi-GPU): this rank's keys that are set, as u64 pairs (read, key), into
// the caller's device buffer (capacity in pairs); returns how many there are -- more than the capacity: nothing usable was
// written, fall back to the dense all-reduce.  apply takes pairs gathered from all ranks (read >= n: padding).
int disco_gpu_compact_keys(disco_ctx *ctx, void *d_pairs, uint64_t capacity, uint64_t *n_pairs)
{
    if (!ctx || !ctx->begun || !d_pairs || !n_pairs) return fail(ctx, DISCO_E_ARG, "compact_keys: bad arguments");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemsetAsync(ctx->d_cursors + CUR_CROWS, 0, sizeof(unsigned long long), ctx->stream)); // (free until finish_contained)
    CK(launch_compact_keys(ctx->d_best, ctx->reads.n, static_cast<unsigned long long *>(d_pairs), capacity, ctx->d_cursors + CUR_CROWS, ctx->stream));
    unsigned long long c = 0;
    CK(cudaMemcpyAsync(&c, ctx->d_cursors + CUR_CROWS, sizeof c, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    *n_pairs = c;
    return DISCO_OK;
}

int disco_gpu_apply_keys(disco_ctx *ctx, const void *d_pairs, uint64_t n_pairs)
{
    if (!ctx || !ctx->begun || (!d_pairs && n_pairs)) return fail(ctx, DISCO_E_ARG, "apply_keys: bad arguments");
    CK(cudaSetDevice(ctx->device));
    CK(launch_apply_keys(ctx->d_best, ctx->reads.n, static_cast<const unsigned long long *>(d_pairs), n_pairs, ctx->stream));
    return DISCO_OK;
}

int disco_gpu_rebase_rows(disco_ctx *ctx, uint64_t u_lo, uint64_t u_hi, uint64_t base)
{
    if (!ctx || !ctx->have_edges) return fail(ctx, DISCO_E_ARG, "edge pass not finished");
    if (u_lo > u_hi || u_hi > ctx->reads.n) return fail(ctx, DISCO_E_ARG, "bad node range");
    CK(cudaSetDevice(ctx->device));
    CK(launch_rebase_rowinfo(ctx->d_rowinfo, u_lo, u_hi, base, ctx->stream));
    return DISCO_OK;
}

int disco_gpu_reserve_rows(disco_ctx *ctx, uint64_t n_entries)
{
    if (!ctx || !ctx->have_edges) return fail(ctx, DISCO_E_ARG, "edge pass not finished");
    CK(cudaSetDevice(ctx->device));
    if (n_entries <= ctx->rows_cap) return DISCO_OK;
    uint64_t *nr = nullptr;
    CK(cudaMalloc(&nr, n_entries * sizeof(uint64_t)));
    if (ctx->rows_used) CK(cudaMemcpyAsync(nr, ctx->d_rows, ctx->rows_used * sizeof(uint64_t), cudaMemcpyDeviceToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    dfree(ctx->d_rows);
    ctx->d_rows = ctx->d_rows_active = nr;
    ctx->rows_cap = n_entries;
    ctx->stats.edge_capacity = n_entries;
    return DISCO_OK;
}

int disco_gpu_move_rows(disco_ctx *ctx, uint64_t dst_offset)
{
    if (!ctx || !ctx->have_edges) return fail(ctx, DISCO_E_ARG, "edge pass not finished");
    if (dst_offset == 0 || ctx->rows_used == 0) return DISCO_OK;
    if (dst_offset < ctx->rows_used || dst_offset + ctx->rows_used > ctx->rows_cap) return fail(ctx, DISCO_E_ARG, "move_rows: bad destination");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(ctx->d_rows + dst_offset, ctx->d_rows, ctx->rows_used * sizeof(uint64_t), cudaMemcpyDeviceToDevice, ctx->stream));
    return DISCO_OK;
}

int disco_gpu_use_rows(disco_ctx *ctx, const uint64_t *d_rows, uint64_t n_entries)
{
    if (!ctx || !ctx->have_edges || !d_rows) return fail(ctx, DISCO_E_ARG, "edge pass not finished");
    (void)n_entries;
    ctx->d_rows_active = const_cast<uint64_t *>(d_rows); // caller-owned; must outlive phase_reduce / get_row
    return DISCO_OK;
}

int disco_gpu_set_rows_used(disco_ctx *ctx, uint64_t n_entries)
{
    if (!ctx || n_entries > ctx->rows_cap) return fail(ctx, DISCO_E_ARG, "set_rows_used: beyond capacity");
    ctx->rows_used = n_entries;
    return DISCO_OK;
}

int disco_gpu_adopt_rows(disco_ctx *ctx, const uint64_t *d_rows, uint64_t n_entries)
{
    if (!ctx || !ctx->have_edges) return fail(ctx, DISCO_E_ARG, "edge pass not finished");
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    dfree(ctx->d_rows);
    ctx->rows_cap = 0;
    CK(cudaMalloc(&ctx->d_rows, std::max<uint64_t>(n_entries, 1) * sizeof(uint64_t)));
    ctx->d_rows_active = ctx->d_rows;
    ctx->rows_cap = ctx->rows_used = n_entries;
    if (n_entries) CK(cudaMemcpyAsync(ctx->d_rows, d_rows, n_entries * sizeof(uint64_t), cudaMemcpyDeviceToDevice, ctx->stream));
    return DISCO_OK;
}

int disco_gpu_set_max_degree(disco_ctx *ctx, uint64_t max_degree)
{
    if (!ctx) return DISCO_E_ARG;
    ctx->stats.max_degree = max_degree;
    return DISCO_OK;
}

int disco_gpu_sync(disco_ctx *ctx)
{
    if (!ctx) return DISCO_E_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    return DISCO_OK;
}

} // extern "C"
