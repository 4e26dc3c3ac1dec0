// kernels.cu -- hand-written sm_100a kernels of the BuildGraph hot path.  Integer / bit work only (hashing, 2-bit
// compares, neighbour-list marking): bound by random 32/64-byte HBM accesses and issued instructions, not by math, so
// no tensor cores.
//
//   k_table_insert_lanes / k_table_insert   HashTable::insertIntoTable (HashTable.cpp:423-514): two records per read,
//                                           quad-cooperative CAS; key-sharded mode inserts only this GPU's keys
//   k_contain_uniform / k_search<CONTAIN>   markContainedReads (OverlapGraph.cpp:333-505) + checkOverlapForContainedRead (:517)
//   k_edges_probe -> k_edges_verify -> k_edges_exact
//                                           insertAllEdgesOfRead (OverlapGraph.cpp:631-678) + checkOverlap (:567);
//                                           k_search<EDGES> is the older fused variant (DISCO_FUSED=1)
//   k_reduce_mark                           markTransitiveEdges (OverlapGraph.cpp:687-723)
//   k_reduce_emit                           removeTransitiveEdges + canonical src<dst selection (:731-761, :808)
//   k_min_keys, k_revcomp_rows, k_restride, k_rebase_rowinfo, k_contained_finish/rows   small helpers
//
// Common shape: one warp (or a 16-lane half of one) per read, the read (forward + reverse complement) staged in shared
// memory, one lane per k-mer position / per candidate so that a warp keeps 32 independent sectors in flight; reads are
// handed out in chunks through an atomic work counter (persistent grid sized from the SM count).  <SHARDED> variants
// read table buckets / adjacency rows of other GPUs through NVLink peer pointers.
#include "dna.cuh"
#include "kernels.cuh"
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include "../../include/disco_gpu.h"

namespace disco {

#define FULL 0xffffffffu
constexpr int kThreads = 256;  // 8 warps per block
constexpr int kWarps = kThreads / 32;
constexpr int kChunk = 16;     // reads per work-counter grab
constexpr int kScanLimit = 6;  // buckets a lane may walk on the fast path before the read is deferred to the slow path
constexpr int kRowBlock = 1024; // adjacency entries a warp reserves per global atomic
constexpr int kHitCap = 128;    // fast-path candidate queue entries per warp (edge pass)
constexpr int kParkMax = 512;    // candidates one read may park for the verify kernel (queue flushed in pieces); more -> exact path
constexpr int kContainQueue = 64; // candidate queue entries per warp (containment pass, drained every 32 positions)
constexpr int kBestMax = 16;   // slow path: smallest-record candidates kept per position (>= 2 * cap)

// ---------------------------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------------------------
// L2 policies: the table and the packed reads are touched at random and never again soon (evict first, and fetch
// 64 instead of 128 bytes of DRAM per miss -- measured in profiles/gather_bench.cu); the presence filter is small and
// hit by every probe (evict last, so that it stays L2 resident while the other two stream through).
__device__ __forceinline__ uint64_t policy_evict_first()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_last()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}

__device__ __forceinline__ void load_bucket(const uint64_t *slots, uint64_t b, uint64_t (&v)[4], uint64_t pol)
{
    // one 32-byte sector, one instruction (LDG.E.256, sm_100+)
    asm volatile("ld.global.nc.L2::cache_hint.L2::64B.v4.u64 {%0,%1,%2,%3}, [%4], %5;"
                 : "=l"(v[0]), "=l"(v[1]), "=l"(v[2]), "=l"(v[3])
                 : "l"(slots + 4 * b), "l"(pol));
}

// pull the adjacency row described by `ri` into L2 (one 128-byte line per lane; rows longer than 512 entries: the head)
__device__ __forceinline__ void prefetch_row(const uint64_t *rows, uint64_t ri, int lane)
{
    const int deg = (int)rowinfo_deg(ri);
    if (lane * 16 < deg) asm volatile("prefetch.global.L2 [%0];" ::"l"(rows + rowinfo_start(ri) + lane * 16));
}

// where the chain of fingerprint h starts: the shard (a peer pointer when the table is key-sharded) and the bucket in it
struct Home {
    const uint64_t *slots;
    uint64_t b;
};
__device__ __forceinline__ uint32_t shard_of(const TableView &t, uint64_t h) { return (uint32_t)__umul64hi(h, (uint64_t)t.world); }
__device__ __forceinline__ uint64_t shard_bucket(const TableView &t, uint64_t h)
{
    // the fraction of h * world left after taking the shard picks the bucket: uniform inside the shard
    return __umul64hi(t.world > 1 ? h * (uint64_t)t.world : h, t.nbuckets);
}
__device__ __forceinline__ Home home_of(const TableView &t, uint64_t h)
{
    Home o;
    o.slots = (t.world > 1) ? t.peers[shard_of(t, h)] : t.slots;
    o.b = shard_bucket(t, h);
    return o;
}

__device__ __forceinline__ bool filter_test(const TableView &t, uint64_t h, uint64_t pol)
{
    if (!t.filter) return true;
    const uint32_t bit = (uint32_t)(h >> 13) & t.filter_mask;
    uint32_t w;
    asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(w) : "l"(t.filter + (bit >> 5)), "l"(pol));
    return (w >> (bit & 31)) & 1;
}

// the same without a cache hint (measured, profiles/filter_bench.cu: evict_last / a persisting window change nothing)
__device__ __forceinline__ bool filter_test(const TableView &t, uint64_t h)
{
    if (!t.filter) return true;
    const uint32_t bit = (uint32_t)(h >> 13) & t.filter_mask;
    return (__ldg(t.filter + (bit >> 5)) >> (bit & 31)) & 1;
}

// candidate read held in registers: NW = words loaded (even, >= words of the longest read); the row stride in memory
// may be larger (power of two)
template <int NW>
struct RegMatcher {
    uint64_t v[NW];
    int stride;
    __device__ __forceinline__ void load(const uint64_t *words, uint64_t r)
    {
        const uint64_t *p = words + r * (uint64_t)stride;
        const uint64_t pol = policy_evict_first();
#pragma unroll
        for (int i = 0; i < NW / 2; i++) // 128-bit loads, rows are 16-byte aligned
            asm volatile("ld.global.nc.L2::cache_hint.L2::64B.v2.u64 {%0,%1}, [%2], %3;"
                         : "=l"(v[2 * i]), "=l"(v[2 * i + 1]) : "l"(p + 2 * i), "l"(pol));
    }
    // bases [a, a+n) of padded array P against bases [b, b+n) of the candidate.  The query window is unaligned by a
    // constant amount for every candidate word, so each word costs two funnel shifts; only the first and last word
    // need a mask.  Branch-free: words outside [wlo, whi] are fetched from a clamped address and masked to zero.
    __device__ __forceinline__ bool operator()(const uint32_t *P, int a, int b, int n, int p_u32) const
    {
        const int q0 = a - b + 32;           // query base (in padded coordinates) facing candidate base 0
        const int i0 = q0 >> 4, sh = (q0 & 15) * 2;
        const int wlo = b >> 5, whi = (b + n - 1) >> 5;
        const uint64_t head = ~0ULL >> (2 * (b & 31));
        const int tb = (b + n) & 31;
        const uint64_t tail = tb ? ~(~0ULL >> (2 * tb)) : ~0ULL;
        uint64_t diff = 0;
#pragma unroll
        for (int w = 0; w < NW; w++) {
            int i = i0 + 2 * w;
            i = max(0, min(i, p_u32 - 3));
            const uint32_t w0 = P[i], w1 = P[i + 1], w2 = P[i + 2];
            uint64_t x = (((uint64_t)fsl32(w1, w0, sh) << 32) | fsl32(w2, w1, sh)) ^ v[w];
            uint64_t m = (w >= wlo && w <= whi) ? ~0ULL : 0ULL;
            if (w == wlo) m &= head;
            if (w == whi) m &= tail;
            diff |= x & m;
        }
        return diff == 0;
    }
};
// long reads: walk the candidate's words in global memory
struct GlobalLoader {
    const uint64_t *p;
    __device__ __forceinline__ uint64_t operator()(int w) const { return __ldg(p + w); }
};
template <>
struct RegMatcher<0> {
    LoaderMatcher<GlobalLoader> m;
    int stride;
    __device__ __forceinline__ void load(const uint64_t *words, uint64_t r) { m.s2.p = words + r * (uint64_t)stride; }
    __device__ __forceinline__ bool operator()(const uint32_t *P, int a, int b, int n, int) const { return m(P, a, b, n); }
};

// Verify-kernel compare: the candidate's words are parked in shared memory, transposed (word w of lane l at
// cw[w * GW + l], GW = lanes per group, conflict-free), so that the compare is a plain loop over exactly the words the overlap touches --
// no per-word range selects, no unrolling over the longest read.
template <int GW>
struct SmemMatcher {
    const uint64_t *cw; // + lane already applied
    __device__ __forceinline__ uint64_t word(int w) const { return cw[w * GW]; }
    __device__ __forceinline__ bool operator()(const uint32_t *P, int a, int b, int n) const
    {
        const int q0 = a - b + 32;           // query base (in padded coordinates) facing candidate base 0
        const int sh = (q0 & 15) * 2;
        const int wlo = b >> 5, whi = (b + n - 1) >> 5;
        const uint32_t *q = P + ((q0 >> 4) + 2 * wlo);
        uint32_t w0 = q[0], w1 = q[1], w2 = q[2];
        uint64_t x = (((uint64_t)fsl32(w1, w0, sh) << 32) | fsl32(w2, w1, sh)) ^ word(wlo);
        x &= ~0ULL >> (2 * (b & 31));                       // first word: bases before b do not count
        const int tb = (b + n) & 31;
        const uint64_t tail = tb ? ~(~0ULL >> (2 * tb)) : ~0ULL;
        if (wlo == whi) return (x & tail) == 0;
        uint64_t diff = x;
        for (int w = wlo + 1; w < whi; w++) {
            q += 2;
            w0 = w2; w1 = q[1]; w2 = q[2];
            diff |= (((uint64_t)fsl32(w1, w0, sh) << 32) | fsl32(w2, w1, sh)) ^ word(w);
        }
        q += 2;
        w0 = w2; w1 = q[1]; w2 = q[2];
        x = (((uint64_t)fsl32(w1, w0, sh) << 32) | fsl32(w2, w1, sh)) ^ word(whi);
        return (diff | (x & tail)) == 0;
    }
};

// stage read r into the warp's padded arrays A (forward) and R (reverse complement); WP = words(max_len) + 2 padded
// words per array (stored as base-ordered 32-bit halves, see dna.cuh)
__device__ __forceinline__ void stage_read(const ReadsView &rv, uint64_t r, int L, uint32_t *A, uint32_t *R, int WP, int lane)
{
    const int W = (L + 31) >> 5;
    const uint64_t *src = rv.words + r * (uint64_t)rv.stride;
    for (int w = lane; w < WP; w += 32) pstore(A, w, (w >= 1 && w <= W) ? __ldg(src + (w - 1)) : 0ULL);
    __syncwarp();
    for (int w = lane; w < WP; w += 32) pstore(R, w, (w >= 1 && w <= W) ? rc_word(A, L, W, w - 1) : 0ULL);
    __syncwarp();
}

// group versions (GW lanes of a warp work on one read; gmask = the group's lanes)
template <int GW>
__device__ __forceinline__ void stage_read_g(const ReadsView &rv, uint64_t r, int L, uint32_t *A, uint32_t *R, int WP, int lane, unsigned gmask)
{
    const int W = (L + 31) >> 5;
    const uint64_t *src = rv.words + r * (uint64_t)rv.stride;
    for (int w = lane; w < WP; w += GW) pstore(A, w, (w >= 1 && w <= W) ? __ldg(src + (w - 1)) : 0ULL);
    __syncwarp(gmask);
    for (int w = lane; w < WP; w += GW) pstore(R, w, (w >= 1 && w <= W) ? rc_word(A, L, W, w - 1) : 0ULL);
    __syncwarp(gmask);
}

// same, with the forward words already in registers (lane w holds word w; reads of up to 32 words = 1024 bases)
__device__ __forceinline__ void stage_read_pre(uint64_t myword, int L, uint32_t *A, uint32_t *R, int WP, int lane)
{
    const int W = (L + 31) >> 5;
    for (int w0 = 0; w0 < WP; w0 += 32) { // uniform trip count: the shuffle needs every lane
        const int w = w0 + lane;
        const uint64_t x = __shfl_sync(FULL, myword, (w - 1) & 31);
        if (w < WP) pstore(A, w, (w >= 1 && w <= W) ? x : 0ULL);
    }
    __syncwarp();
    for (int w = lane; w < WP; w += 32) pstore(R, w, (w >= 1 && w <= W) ? rc_word(A, L, W, w - 1) : 0ULL);
    __syncwarp();
}

__device__ __forceinline__ bool is_contained(const uint32_t *bits, uint32_t r) { return (__ldg(bits + (r >> 5)) >> (r & 31)) & 1; }

__device__ __forceinline__ int read_len(const ReadsView &rv, uint64_t r)
{
    return rv.uniform_len ? rv.uniform_len : (int)__ldg(rv.len + r);
}

__device__ __forceinline__ bool grab_chunk(unsigned long long *counter, uint64_t lo, uint64_t hi, int lane,
                                           uint64_t *begin, uint64_t *end)
{
    unsigned long long c = 0;
    if (lane == 0) c = atomicAdd(counter, (unsigned long long)kChunk);
    c = __shfl_sync(FULL, c, 0);
    uint64_t b = lo + c;
    if (b >= hi) return false;
    *begin = b;
    *end = (b + kChunk < hi) ? b + kChunk : hi;
    return true;
}

template <int GW>
__device__ __forceinline__ bool grab_chunk_g(unsigned long long *counter, uint64_t lo, uint64_t hi, int lane, unsigned gmask,
                                             uint64_t *begin, uint64_t *end)
{
    unsigned long long c = 0;
    if (lane == 0) c = atomicAdd(counter, (unsigned long long)kChunk);
    c = __shfl_sync(gmask, c, 0, GW);
    uint64_t b = lo + c;
    if (b >= hi) return false;
    *begin = b;
    *end = (b + kChunk < hi) ? b + kChunk : hi;
    return true;
}

__device__ __forceinline__ void warp_stat_add(unsigned long long *stats, int slot, unsigned long long v)
{
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(stats + slot, v);
}

// ---------------------------------------------------------------------------------------------------------------
// hash table build: two records per read (prefix k-mer, suffix k-mer), bucket = one 32-byte sector of four slots.
// Warp per read; quad 0 inserts the prefix record and quad 1 the suffix record cooperatively: each lane of the quad
// reads one slot of the bucket (one coalesced sector), the quad votes, the first empty lane does the CAS.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_table_insert(ReadsView rv, TableView tv, int K, const uint32_t *skip_bits, uint64_t r_lo, uint64_t r_hi, const unsigned int *gate)
{
    extern __shared__ uint64_t smem[];
    if (gate && !*gate) return;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int WP = ((rv.max_len + 31) >> 5) + 2;
    uint32_t *A = reinterpret_cast<uint32_t *>(smem + (size_t)wib * 2 * WP), *R = A + 2 * WP;
    const int wpb = blockDim.x >> 5;
    const uint64_t nwarps = (uint64_t)gridDim.x * wpb;
    for (uint64_t r = r_lo + (uint64_t)blockIdx.x * wpb + wib; r < r_hi; r += nwarps) {
        if (skip_bits && ((__ldg(skip_bits + (r >> 5)) >> (r & 31)) & 1)) continue; // warp-uniform
        const int L = read_len(rv, r);
        stage_read(rv, r, L, A, R, WP, lane);
        if (lane < 8) {
            const int quad = lane >> 2, sub = lane & 3;
            const unsigned qmask = 0xFu << (quad * 4);
            int fwd;
            // record 2r = prefix k-mer (j = 0), record 2r+1 = suffix k-mer (j = L-K): HashTable.cpp:430-431
            const uint64_t h = canon_kmer_hash(A, R, L, quad ? L - K : 0, K, &fwd);
            const uint64_t val = make_slot(h, fwd, (uint32_t)(2 * r + quad));
            if (tv.filter && sub == 0) {
                const uint32_t bit = (uint32_t)(h >> 13) & tv.filter_mask;
                atomicOr(tv.filter + (bit >> 5), 1u << (bit & 31));
            }
            const bool mine = !(tv.world > 1 && shard_of(tv, h) != tv.rank); // else another GPU's key (uniform within the quad)
            uint64_t b = shard_bucket(tv, h);
            for (uint64_t walked = 0; mine; ) {
                unsigned long long *slot = reinterpret_cast<unsigned long long *>(tv.slots) + 4 * b + sub;
                const uint64_t cur = __ldcg(slot); // L2 (coherent) read: slots change under our feet
                const unsigned empties = (__ballot_sync(qmask, cur == kEmptySlot) >> (quad * 4)) & 0xFu;
                if (empties) {
                    const int first = __ffs(empties) - 1;
                    int ok = 0;
                    if (sub == first) ok = atomicCAS(slot, (unsigned long long)kEmptySlot, (unsigned long long)val) == kEmptySlot;
                    ok = __shfl_sync(qmask, ok, quad * 4 + first);
                    if (ok) break; // otherwise somebody else took it: vote again on the same bucket
                } else {
                    if (++walked == tv.nbuckets) { if (sub == 0) atomicExch(tv.full, 1u); break; } // (shard) full: reported, not spun on
                    b = (b + 1 == tv.nbuckets) ? 0 : b + 1;
                }
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Lane-per-read variants (short reads): 32 reads per warp step, each lane owns private padded arrays in shared memory
// (odd 32-bit stride -> conflict-free), so hashing runs with all 32 lanes busy.
// ---------------------------------------------------------------------------------------------------------------
__host__ __device__ inline int lane_array_u32(int max_len) { return 2 * (((max_len + 31) >> 5) + 2) + 1; } // one padded array, odd

// stage this lane's read r (or nothing when !valid) into its private arrays
__device__ __forceinline__ void stage_lane(const ReadsView &rv, uint64_t r, bool valid, int L, uint32_t *A, uint32_t *R, int WP)
{
    const int W = (L + 31) >> 5;
    if (valid) {
        const uint64_t *src = rv.words + r * (uint64_t)rv.stride;
        pstore(A, 0, 0ULL); pstore(R, 0, 0ULL);
        for (int w = 1; w <= W; w++) pstore(A, w, __ldg(src + (w - 1)));
        for (int w = W + 1; w < WP; w++) { pstore(A, w, 0ULL); pstore(R, w, 0ULL); }
        for (int w = 1; w <= W; w++) pstore(R, w, rc_word(A, L, W, w - 1));
    }
}

// Table build, 32 reads per warp step.  Each lane hashes the prefix and suffix k-mer of its read; the 64 records are
// then inserted eight at a time, one record per quad, cooperatively (four lanes read the four slots of the bucket --
// one sector --, vote, the first empty lane does the CAS).
__global__ void __launch_bounds__(kThreads) k_table_insert_lanes(ReadsView rv, TableView tv, int K, const uint32_t *skip_bits, uint64_t r_lo, uint64_t r_hi, const unsigned int *gate)
{
    extern __shared__ uint64_t smem[];
    if (gate && !*gate) return;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int WP = ((rv.max_len + 31) >> 5) + 2;
    const int AU = lane_array_u32(rv.max_len);
    uint32_t *base = reinterpret_cast<uint32_t *>(smem) + (size_t)wib * 64 * AU;
    uint32_t *A = base + (size_t)lane * 2 * AU, *R = A + AU;
    const int wpb = blockDim.x >> 5;
    const uint64_t nwarps = (uint64_t)gridDim.x * wpb;
    const int quad = lane >> 2, sub = lane & 3;
    const unsigned qmask = 0xFu << (quad * 4);
    for (uint64_t r0 = r_lo + ((uint64_t)blockIdx.x * wpb + wib) * 32; r0 < r_hi; r0 += nwarps * 32) {
        const uint64_t r = r0 + lane;
        bool valid = r < r_hi;
        if (valid && skip_bits) valid = !((__ldg(skip_bits + (r >> 5)) >> (r & 31)) & 1);
        const int L = valid ? read_len(rv, r) : 0;
        stage_lane(rv, r, valid, L, A, R, WP);
        uint64_t h0 = 0, h1 = 0;
        int f0 = 0, f1 = 0;
        if (valid) {
            // record 2r = prefix k-mer (j = 0), record 2r+1 = suffix k-mer (j = L-K): HashTable.cpp:430-431
            h0 = canon_kmer_hash(A, R, L, 0, K, &f0);
            h1 = canon_kmer_hash(A, R, L, L - K, K, &f1);
            if (tv.filter) {
                uint32_t bit = (uint32_t)(h0 >> 13) & tv.filter_mask;
                atomicOr(tv.filter + (bit >> 5), 1u << (bit & 31));
                bit = (uint32_t)(h1 >> 13) & tv.filter_mask;
                atomicOr(tv.filter + (bit >> 5), 1u << (bit & 31));
            }
        }
        const uint64_t v0 = make_slot(h0, f0, (uint32_t)(2 * r)), v1 = make_slot(h1, f1, (uint32_t)(2 * r + 1));
        for (int t = 0; t < 8; t++) { // records t*8 .. t*8+7, one per quad
            const int rec = t * 8 + quad, owner = rec >> 1;
            // (select after the shuffle: the source lane would evaluate the selector with its own record index)
            const uint64_t ha = __shfl_sync(FULL, h0, owner), hb = __shfl_sync(FULL, h1, owner);
            const uint64_t va = __shfl_sync(FULL, v0, owner), vb = __shfl_sync(FULL, v1, owner);
            const uint64_t hh = (rec & 1) ? hb : ha, val = (rec & 1) ? vb : va;
            const int ok_rec = __shfl_sync(FULL, (int)valid, owner);
            if (!ok_rec) continue; // uniform within the quad
            if (tv.world > 1 && shard_of(tv, hh) != tv.rank) continue; // another GPU's key
            uint64_t b = shard_bucket(tv, hh);
            for (uint64_t walked = 0;;) {
                unsigned long long *slot = reinterpret_cast<unsigned long long *>(tv.slots) + 4 * b + sub;
                const uint64_t cur = __ldcg(slot); // L2 (coherent) read: slots change under our feet
                const unsigned empties = (__ballot_sync(qmask, cur == kEmptySlot) >> (quad * 4)) & 0xFu;
                if (empties) {
                    const int first = __ffs(empties) - 1;
                    int ok = 0;
                    if (sub == first) ok = atomicCAS(slot, (unsigned long long)kEmptySlot, (unsigned long long)val) == kEmptySlot;
                    ok = __shfl_sync(qmask, ok, quad * 4 + first);
                    if (ok) break; // otherwise somebody else took it: vote again on the same bucket
                } else {
                    if (++walked == tv.nbuckets) { if (sub == 0) atomicExch(tv.full, 1u); break; } // (shard) full: reported, not spun on
                    b = (b + 1 == tv.nbuckets) ? 0 : b + 1;
                }
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Binned table build.  Inserting 2n records at random into a table of 96n bytes costs two DRAM transactions per record
// (the bucket's sector comes in, the dirty sector goes out) at the random-access rate -- and on a table replicated over
// the GPUs of a node every rank inserts ALL reads.  So the records are first sorted by bucket range: k_table_bin hashes
// the reads (lane per read, as k_table_insert_lanes) and appends every (fingerprint, slot value) to the bin its bucket
// falls into -- a block-wide histogram in shared memory, one global atomic per bin and block step, 16-byte stores that L2
// merges into full lines; k_table_fill then walks the bins in order with a few warps per SM, so that the 16 MB slice of
// the table one bin maps to is filled while it sits in L2: the inserts become L2 atomics and the slice goes to DRAM once.
// The table's content is the same set of records; which of a bucket's four slots a record lands in was never defined.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_table_bin(ReadsView rv, TableView tv, int K, BinView bv, uint64_t r_lo, uint64_t r_hi)
{
    extern __shared__ uint64_t smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int WP = ((rv.max_len + 31) >> 5) + 2;
    const int AU = lane_array_u32(rv.max_len);
    uint32_t *base = reinterpret_cast<uint32_t *>(smem) + (size_t)wib * 64 * AU;
    uint32_t *A = base + (size_t)lane * 2 * AU, *R = A + AU;
    // behind the lane arrays: where this step's records of each bin start (u64[nbins]) and how many there are (u32[nbins])
    unsigned long long *start = reinterpret_cast<unsigned long long *>(smem + ((size_t)kWarps * 64 * AU + 1) / 2);
    uint32_t *hist = reinterpret_cast<uint32_t *>(start + bv.nbins);
    const uint32_t nbins = bv.nbins;
    for (uint64_t rb = r_lo + (uint64_t)blockIdx.x * kThreads; rb < r_hi; rb += (uint64_t)gridDim.x * kThreads) { // block-uniform
        const uint64_t r = rb + threadIdx.x;
        const bool valid = r < r_hi;
        const int L = valid ? read_len(rv, r) : 0;
        stage_lane(rv, r, valid, L, A, R, WP);
        uint64_t h0 = 0, h1 = 0;
        int f0 = 0, f1 = 0;
        if (valid) {
            // record 2r = prefix k-mer (j = 0), record 2r+1 = suffix k-mer (j = L-K): HashTable.cpp:430-431
            h0 = canon_kmer_hash(A, R, L, 0, K, &f0);
            h1 = canon_kmer_hash(A, R, L, L - K, K, &f1);
            if (tv.filter) {
                uint32_t bit = (uint32_t)(h0 >> 13) & tv.filter_mask;
                atomicOr(tv.filter + (bit >> 5), 1u << (bit & 31));
                bit = (uint32_t)(h1 >> 13) & tv.filter_mask;
                atomicOr(tv.filter + (bit >> 5), 1u << (bit & 31));
            }
        }
        if (nbins == 1) { // no sorting: record i of the table at recs[i] (32 contiguous bytes per lane)
            if (valid) {
                bv.recs[2 * r] = make_ulonglong2(h0, make_slot(h0, f0, (uint32_t)(2 * r)));
                bv.recs[2 * r + 1] = make_ulonglong2(h1, make_slot(h1, f1, (uint32_t)(2 * r + 1)));
            }
            continue;
        }
        for (uint32_t t = threadIdx.x; t < nbins; t += kThreads) hist[t] = 0;
        __syncthreads();
        // bucket = mulhi(h, nbuckets) and bin = mulhi(h, nbins) are both monotone in h: a bin is a range of buckets
        const uint32_t b0 = (uint32_t)__umul64hi(h0, (uint64_t)nbins), b1 = (uint32_t)__umul64hi(h1, (uint64_t)nbins);
        uint32_t k0 = 0, k1 = 0;
        if (valid) { k0 = atomicAdd(&hist[b0], 1u); k1 = atomicAdd(&hist[b1], 1u); }
        __syncthreads();
        for (uint32_t t = threadIdx.x; t < nbins; t += kThreads) {
            const uint32_t c = hist[t];
            start[t] = c ? atomicAdd(bv.count + t, (unsigned long long)c) : 0ULL;
        }
        __syncthreads();
        if (valid) {
            const unsigned long long p0 = start[b0] + k0, p1 = start[b1] + k1;
            if (p0 < bv.cap) bv.recs[(uint64_t)b0 * bv.cap + p0] = make_ulonglong2(h0, make_slot(h0, f0, (uint32_t)(2 * r)));
            else atomicExch(bv.overflow, 1u);
            if (p1 < bv.cap) bv.recs[(uint64_t)b1 * bv.cap + p1] = make_ulonglong2(h1, make_slot(h1, f1, (uint32_t)(2 * r + 1)));
            else atomicExch(bv.overflow, 1u);
        }
        // (the next step's writes to hist / start are separated from this step's reads by its first two barriers)
    }
}

// k_table_fill: one lane per record -- 32 independent bucket loads + CAS per warp instead of the eight a warp of the direct
// kernel has in flight between its hashing steps.  (nbins == 1: the records lie in read order, record i at recs[i].)
__global__ void __launch_bounds__(kThreads) k_table_fill(TableView tv, BinView bv, const uint32_t *skip_bits, int set_filter,
                                                         unsigned long long *work, uint32_t chunks_per_bin, uint32_t fill_chunk,
                                                         unsigned long long n_unbinned)
{
    const int lane = threadIdx.x & 31;
    if (*reinterpret_cast<volatile unsigned int *>(bv.overflow)) return; // incomplete bins: the gated direct kernel builds the table
    const unsigned long long items = (unsigned long long)bv.nbins * chunks_per_bin;
    unsigned long long *slots = reinterpret_cast<unsigned long long *>(tv.slots);
    for (;;) {
        unsigned long long it = 0;
        if (lane == 0) it = atomicAdd(work, 1ULL);
        it = __shfl_sync(FULL, it, 0);
        if (it >= items) break;
        const uint32_t b = (uint32_t)(it / chunks_per_bin), c = (uint32_t)(it % chunks_per_bin);
        const unsigned long long cnt = bv.nbins == 1 ? n_unbinned : __ldg(bv.count + b); // (final: k_table_bin has finished)
        const unsigned long long lo = (unsigned long long)c * fill_chunk;
        if (lo >= cnt) continue;
        const unsigned long long hi = cnt < lo + fill_chunk ? cnt : lo + fill_chunk;
        const ulonglong2 *recs = bv.recs + (uint64_t)b * bv.cap;
        for (unsigned long long i = lo + lane; i < hi; i += 32) {
            const ulonglong2 rec = __ldcs(recs + i);
            const uint64_t h = rec.x, val = rec.y;
            if (val == kEmptySlot) continue; // (unbinned layout: no record -- a read the binning pass did not see)
            if (skip_bits && is_contained(skip_bits, (uint32_t)val >> 1)) continue;
            if (set_filter && tv.filter) {
                const uint32_t bit = (uint32_t)(h >> 13) & tv.filter_mask;
                atomicOr(tv.filter + (bit >> 5), 1u << (bit & 31));
            }
            uint64_t bk = bucket_of(h, tv.nbuckets);
            for (uint64_t walked = 0;;) {
                unsigned long long *s = slots + 4 * bk;
                const ulonglong2 x = __ldcg(reinterpret_cast<const ulonglong2 *>(s)), y = __ldcg(reinterpret_cast<const ulonglong2 *>(s + 2));
                const int q = x.x == kEmptySlot ? 0 : x.y == kEmptySlot ? 1 : y.x == kEmptySlot ? 2 : y.y == kEmptySlot ? 3 : -1;
                if (q >= 0) {
                    if (atomicCAS(s + q, (unsigned long long)kEmptySlot, (unsigned long long)val) == kEmptySlot) break;
                    continue; // somebody else took that slot: look at the bucket again
                }
                if (++walked == tv.nbuckets) { atomicExch(tv.full, 1u); break; } // table full: reported, not spun on
                bk = (bk + 1 == tv.nbuckets) ? 0 : bk + 1;
            }
        }
    }
}

// Containment pass when every read has the same length L: the only feasible position is j = 0 (types 0/2 need
// j + L <= L; types 1/3 need j >= L - K, outside [0, L-K)), i.e. contained == exact duplicate, forward or reverse
// complement, of an earlier read (OverlapGraph.cpp:449).  One probe per read, so a lane per read; the compare is
// word-aligned (whole read against the candidate).
template <int NW>
__global__ void __launch_bounds__(kThreads) k_contain_uniform(SearchParams p)
{
    extern __shared__ uint64_t smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int L = p.reads.uniform_len, K = p.K;
    const int W = (L + 31) >> 5, WP = W + 2;
    const int AU = lane_array_u32(p.reads.max_len);
    uint32_t *base = reinterpret_cast<uint32_t *>(smem) + (size_t)wib * 64 * AU;
    uint32_t *A = base + (size_t)lane * 2 * AU, *R = A + AU;
    const uint64_t pol_stream = policy_evict_first(), pol_keep = policy_evict_last();
    const int wpb = blockDim.x >> 5;
    const uint64_t nwarps = (uint64_t)gridDim.x * wpb;
    unsigned n_queries = 0, n_probes = 0, n_buckets = 0, n_verified = 0, n_hits = 0;
    for (uint64_t r0 = p.q_lo + ((uint64_t)blockIdx.x * wpb + wib) * 32; r0 < p.q_hi; r0 += nwarps * 32) {
        const uint64_t r1 = r0 + lane;
        const bool valid = r1 < p.q_hi;
        stage_lane(p.reads, r1, valid, L, A, R, WP);
        if (!valid) continue;
        n_queries++; n_probes++;
        int fq;
        const uint64_t h = canon_kmer_hash(A, R, L, 0, K, &fq);
        if (!filter_test(p.table, h, pol_keep)) continue; // cannot happen for a read in the table; kept for symmetry
        const uint32_t tag = slot_tag(h);
        const Home home = home_of(p.table, h);
        uint64_t b = home.b;
        for (;;) {
            uint64_t v[4];
            load_bucket(home.slots, b, v, pol_stream);
            n_buckets++;
            bool hole = false;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                if (v[q] == kEmptySlot) { hole = true; continue; }
                if ((uint32_t)(v[q] >> 33) != tag) continue;
                const uint32_t rec = (uint32_t)v[q], r2 = rec >> 1;
                // equal lengths: read1 covers read2 only when it comes first in the file (OverlapGraph.cpp:449)
                if (!(r1 < r2)) continue;
                const int type = cand_type(rec & 1, (int)((v[q] >> 32) & 1) == fq);
                if (type == 1 || type == 3) continue; // suffix-anchored types need j >= L - K
                RegMatcher<NW> m;
                m.stride = p.reads.stride;
                m.load(p.reads.words, r2);
                n_verified++;
                const uint32_t *S = (type == 0) ? A : R; // type 0: s1 == s2 ; type 2: rc(s1) == s2
                uint64_t diff = 0;
#pragma unroll
                for (int w = 0; w < NW; w++)
                    if (w < W) diff |= pword(S, w + 1) ^ m.v[w];
                if (diff == 0) {
                    n_hits++;
                    atomicMin(p.best + r2, (unsigned long long)make_ckey(r1, 0, rec & 1, type));
                }
            }
            if (hole) break;
            b = (b + 1 == p.table.nbuckets) ? 0 : b + 1;
        }
    }
    warp_stat_add(p.stats, ST_QUERIES, n_queries);
    warp_stat_add(p.stats, ST_PROBES, n_probes);
    warp_stat_add(p.stats, ST_BUCKETS, n_buckets);
    warp_stat_add(p.stats, ST_VERIFIED, n_verified);
    warp_stat_add(p.stats, ST_HITS, n_hits);
}

// ---------------------------------------------------------------------------------------------------------------
// search kernels
// ---------------------------------------------------------------------------------------------------------------
enum { MODE_CONTAIN = 0, MODE_EDGES = 1 };

// u64 words of shared memory one warp of k_search needs: A, R (WP each), position list (npos hashes + npos u32),
// candidate queue, and for the edge pass the row buffer, the slow path's best list and 4 control ints
__host__ __device__ inline size_t search_words_per_warp(int WP, int npos, int hcap, int rowcap, int mode)
{
    size_t w = 2 * (size_t)WP + (size_t)npos + ((size_t)npos + 1) / 2 + (size_t)hcap + 2;
    if (mode == MODE_EDGES) w += (size_t)rowcap + kBestMax;
    return w;
}

__host__ __device__ inline int reduce_mark_hset(int maxdeg) { int v = 64; while (v < 2 * maxdeg) v <<= 1; return v; }
__host__ __device__ inline size_t reduce_mark_smem_per_warp(int maxdeg)
{
    return (((size_t)maxdeg * 9 + (size_t)reduce_mark_hset(maxdeg) * 4 + 15) / 16) * 16;
}

// hit record of the edge pass: [55..40 j][39..8 rec][1..0 type]  -> sorts by (j, rec) = the reference's visiting order
__device__ __forceinline__ uint64_t make_hit(int j, uint32_t rec, int type) { return ((uint64_t)j << 40) | ((uint64_t)rec << 8) | (uint64_t)type; }
__device__ __forceinline__ int hit_j(uint64_t h) { return (int)(h >> 40); }
__device__ __forceinline__ uint32_t hit_read(uint64_t h) { return (uint32_t)(h >> 9) & 0x7FFFFFFFu; }
__device__ __forceinline__ int hit_type(uint64_t h) { return (int)(h & 3); }

struct WarpSmem {
    uint32_t *A, *R;
    int p_u32; // 32-bit words in each padded array
    uint64_t *ph;   // hashes of the positions that passed the presence filter
    uint32_t *pj;   // their (position << 1 | canonical-is-forward)
    uint64_t *hits, *row, *best;
    int *ctrl; // [0] n hits, [1] slow flag, [2] n row, [3] n best
};

template <int NW>
__device__ __forceinline__ bool verify_dovetail(const SearchParams &p, const WarpSmem &s, int L1, int j, int type, uint32_t r2)
{
    RegMatcher<NW> m;
    m.stride = p.reads.stride;
    m.load(p.reads.words, r2);
    const int L2 = read_len(p.reads, r2);
    int use_rc, a, b, n;
    if (!dovetail_window(type, L1, j, p.K, L2, &use_rc, &a, &b, &n)) return false;
    return m(use_rc ? s.R : s.A, a, b, n, s.p_u32);
}

// Exact sequential search of one read (used when MAX_EDGE_PER_KMER can fire or a position has a long candidate list):
// positions in ascending order, candidates in record order, at most `cap` insertions per position, a read already in
// the row is skipped without counting (OverlapGraph.cpp:645-670).  The warp walks 8 buckets (32 slots) per step.
template <int NW>
__device__ void search_edges_slow(const SearchParams &p, const WarpSmem &s, uint64_t r1, int L1, int lane,
                                  unsigned &n_probes, unsigned &n_buckets, unsigned &n_verified, unsigned &n_capfired)
{
    const int K = p.K;
    int nrow = 0;
    if (lane == 0) s.ctrl[2] = 0;
    __syncwarp();
    for (int j = 1; j < L1 - K; j++) {
        int fq;
        const uint64_t h = canon_kmer_hash(s.A, s.R, L1, j, K, &fq);
        const uint32_t tag = slot_tag(h);
        const Home home = home_of(p.table, h);
        uint64_t b = home.b;
        int nbest = 0, nvalid = 0;
        n_probes += (lane == 0);
        for (;;) {
            uint64_t bb = b + (lane >> 2);
            if (bb >= p.table.nbuckets) bb -= p.table.nbuckets;
            const uint64_t v = __ldg(home.slots + 4 * bb + (lane & 3));
            const unsigned empties = __ballot_sync(FULL, v == kEmptySlot);
            const int limit = empties ? ((__ffs(empties) - 1) | 3) : 31; // the chain ends in the first bucket with a hole
            n_buckets += (lane == 0) ? (unsigned)((limit >> 2) + 1) : 0u;
            bool valid = false;
            uint64_t key = 0;
            if (lane <= limit && v != kEmptySlot && (uint32_t)(v >> 33) == tag) {
                const uint32_t rec = (uint32_t)v, r2 = rec >> 1;
                bool seen = (r2 == (uint32_t)r1) || (p.skip_contained && is_contained(p.contained_bits, r2));
                for (int k = 0; k < nrow && !seen; k++) seen = (uint32_t)entry_nbr(s.row[k]) == r2;
                if (!seen) {
                    const int type = cand_type(rec & 1, (int)((v >> 32) & 1) == fq);
                    n_verified++;
                    if (verify_dovetail<NW>(p, s, L1, j, type, r2)) { valid = true; key = ((uint64_t)rec << 2) | (uint64_t)type; }
                }
            }
            unsigned vm = __ballot_sync(FULL, valid);
            while (vm) { // keep the kBestMax smallest records, sorted (lane 0 owns the list)
                const int src = __ffs(vm) - 1; vm &= vm - 1;
                const uint64_t kk = __shfl_sync(FULL, key, src);
                nvalid++;
                if (lane == 0) {
                    int pos = nbest;
                    if (nbest == kBestMax) { if (kk >= s.best[nbest - 1]) pos = -1; else pos = nbest - 1; }
                    else nbest++;
                    if (pos >= 0) {
                        while (pos > 0 && s.best[pos - 1] > kk) { s.best[pos] = s.best[pos - 1]; pos--; }
                        s.best[pos] = kk;
                    }
                }
            }
            if (empties) break;
            b += 8; if (b >= p.table.nbuckets) b -= p.table.nbuckets;
        }
        if (lane == 0) {
            int ctr = 0, fired = 0;
            const int row0 = nrow;
            for (int t = 0; t < nbest; t++) {
                const uint64_t kk = s.best[t];
                const uint32_t r2 = (uint32_t)(kk >> 3);
                bool seen = false;
                for (int k = row0; k < nrow && !seen; k++) seen = (uint32_t)entry_nbr(s.row[k]) == r2; // same position, other record
                if (seen) continue;
                if (ctr >= p.cap) { fired = 1; continue; }
                int orient, ovl;
                type_to_edge((int)(kk & 3), L1, K, j, &orient, &ovl);
                if (nrow < p.rowcap) s.row[nrow] = make_entry(L1 - ovl, r2, orient);
                nrow++; ctr++;
            }
            if (nvalid > nbest && ctr >= p.cap) fired = 1;
            n_capfired += fired;
            s.ctrl[2] = nrow;
        }
        __syncwarp();
        nrow = s.ctrl[2];
    }
}

// containment test of one queued candidate + the atomicMin that elects the container (OverlapGraph.cpp:421-449)
template <int NW>
__device__ __forceinline__ bool contain_one(const SearchParams &p, const WarpSmem &s, uint64_t r1, int L1, uint64_t c)
{
    const uint32_t r2 = hit_read(c);
    const int L2 = read_len(p.reads, r2);
    // read1 must be longer, or equal and earlier in the file (OverlapGraph.cpp:424, :449)
    if (!(L1 > L2 || (L1 == L2 && r1 < r2))) return false;
    RegMatcher<NW> m;
    m.stride = p.reads.stride;
    m.load(p.reads.words, r2);
    const int j = hit_j(c), type = hit_type(c);
    int use_rc, a, b, n;
    if (!contained_window(type, L1, j, p.K, L2, &use_rc, &a, &b, &n)) return false;
    if (!m(use_rc ? s.R : s.A, a, b, n, s.p_u32)) return false;
    atomicMin(p.best + r2, (unsigned long long)make_ckey(r1, j, (int)((c >> 8) & 1), type));
    return true;
}

// Search kernel, both passes.  Per read (one warp):
//   1a. hash   : one lane per k-mer position: canonical fingerprint, presence-filter bit (L2 resident); positions that
//                pass are ballot-compacted into a list and their table bucket is prefetched into L2
//   1b. probe  : one lane per listed position (lanes packed): one 32-byte bucket load, branch-free tag compare; every
//                tag match is queued as (position, record, type) and the candidate read is prefetched
//   2. verify  : one lane per queued candidate: 128-bit loads of the candidate read, one windowed 2-bit compare
//                against the staged query (forward or reverse complement)
//   3. resolve : (edge pass) first hit per neighbour wins (OverlapGraph.cpp:656) -- checked with a shared-memory
//                set of neighbour ids; positions with more than `cap` partners send the read to the exact
//                sequential path; survivors are compacted by ballot and appended to the adjacency
template <int NW, int MODE>
__global__ void __launch_bounds__(kThreads, (NW > 0 && NW <= 8) ? 4 : 1) k_search(SearchParams p)
{
    extern __shared__ uint64_t smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int WP = ((p.reads.max_len + 31) >> 5) + 2;
    const int K = p.K;
    // per-warp shared memory carve-up (u64 units); must match search_smem_per_warp()
    const size_t per_warp = search_words_per_warp(WP, p.npos, p.hcap, p.rowcap, MODE);
    WarpSmem s;
    uint64_t *w0 = smem + wib * per_warp;
    s.A = reinterpret_cast<uint32_t *>(w0); s.R = s.A + 2 * WP; s.p_u32 = 2 * WP;
    s.ph = w0 + 2 * WP;
    s.pj = reinterpret_cast<uint32_t *>(s.ph + p.npos);
    s.hits = s.ph + p.npos + (p.npos + 1) / 2;
    if (MODE == MODE_EDGES) { s.row = s.hits + p.hcap; s.best = s.row + p.rowcap; s.ctrl = reinterpret_cast<int *>(s.best + kBestMax); }
    else { s.row = nullptr; s.best = nullptr; s.ctrl = reinterpret_cast<int *>(s.hits + p.hcap); }
    // fast-path scratch inside the (otherwise idle) row buffer: neighbour-id set + per-position counters
    uint32_t *hset = (MODE == MODE_EDGES) ? reinterpret_cast<uint32_t *>(s.row) : nullptr;
    int *cntj = (MODE == MODE_EDGES) ? reinterpret_cast<int *>(s.row) + p.hset : nullptr;
    const unsigned lt_mask = (1u << lane) - 1;
    const uint64_t nbuckets = p.table.nbuckets;
    const int hcap = p.hcap, cap = p.cap;

    unsigned n_queries = 0, n_probes = 0, n_buckets = 0, n_verified = 0, n_hits = 0, n_capfired = 0, n_slow = 0, maxdeg = 0;
    unsigned long long n_entries = 0;
    unsigned long long blk_cur = 0, blk_end = 0; // this warp's reserved slice of the adjacency buffer
    const bool use_pre = p.reads.stride <= 32;
    uint64_t pre_for = ~0ULL, pre_word = 0;
    const uint64_t pol_stream = policy_evict_first(), pol_keep = policy_evict_last();
    uint64_t rb, re;
    while (grab_chunk(p.work_counter, p.q_lo, p.q_hi, lane, &rb, &re)) {
        for (uint64_t r1 = rb; r1 < re; r1++) {
            if (MODE == MODE_EDGES && ((__ldg(p.contained_bits + (r1 >> 5)) >> (r1 & 31)) & 1)) continue; // OverlapGraph.cpp:657
            if (MODE == MODE_CONTAIN && p.only_flagged && !(p.rowinfo[r1] & kInfoExact)) continue;       // fall-back of the flat pass
            const int L1 = read_len(p.reads, r1);
            if (use_pre) {
                // the words of this read were requested while the previous one was processed
                const uint64_t mine = (pre_for == r1) ? pre_word : ((lane < p.reads.stride) ? __ldg(p.reads.words + r1 * (uint64_t)p.reads.stride + lane) : 0ULL);
                if (r1 + 1 < re) { pre_for = r1 + 1; pre_word = (lane < p.reads.stride) ? __ldg(p.reads.words + (r1 + 1) * (uint64_t)p.reads.stride + lane) : 0ULL; }
                stage_read_pre(mine, L1, s.A, s.R, WP, lane);
            } else {
                stage_read(p.reads, r1, L1, s.A, s.R, WP, lane);
            }
            n_queries += (lane == 0);
            if (lane < 4) s.ctrl[lane] = 0; // [0] queued, [1] needs exact path, [2] row length, [3] some position has > cap candidates
            __syncwarp();
            // containment: positions [0, L1-K) (OverlapGraph.cpp:401), pruned to those where some read can fit:
            //   types 0/2 need j + L2 <= L1, types 1/3 need j >= L2 - K, and L2 >= min_len.
            // edges: positions [1, L1-K) (OverlapGraph.cpp:638)
            const int jlo = (MODE == MODE_EDGES) ? 1 : 0, jhi = L1 - K;
            // ---- 1a. hash + presence filter -------------------------------------------------------------------
            int np = 0;
            for (int jb = jlo; jb < jhi; jb += 32) {
                const int j = jb + lane;
                bool act = j < jhi;
                if (MODE == MODE_CONTAIN) act = act && (j <= L1 - p.reads.min_len || j >= p.reads.min_len - K);
                uint64_t h = 0;
                int fq = 0;
                bool pass = false;
                if (act) {
                    h = canon_kmer_hash(s.A, s.R, L1, j, K, &fq);
                    n_probes++;
                    // most positions match no record at all: the L2-resident presence bit answers that without DRAM
                    pass = filter_test(p.table, h, pol_keep);
                    if (pass && p.table.world <= 1) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.table.slots + 4 * bucket_of(h, nbuckets)));
                }
                const unsigned m = __ballot_sync(FULL, pass);
                if (pass) {
                    const int pos = np + __popc(m & lt_mask);
                    s.ph[pos] = h;
                    s.pj[pos] = ((uint32_t)j << 1) | (uint32_t)fq;
                }
                np += __popc(m);
            }
            __syncwarp();
            if (p.dbg & 2) continue;
            // ---- 1b. probe: one lane per surviving position -----------------------------------------------------
            for (int i0 = 0; i0 < np; i0 += 32) {
                const int i = i0 + lane;
                if (i < np) {
                    const uint64_t h = s.ph[i];
                    const uint32_t jf = s.pj[i];
                    const int j = (int)(jf >> 1), fq = (int)(jf & 1);
                    const uint32_t tag = slot_tag(h);
                    const Home home = home_of(p.table, h);
                    uint64_t b = home.b;
                    int pushed = 0;
                    for (int walked = 0;; walked++) {
                        if (MODE == MODE_EDGES && walked == kScanLimit) { s.ctrl[1] = 1; break; } // long chain: exact path
                        uint64_t v[4];
                        load_bucket(home.slots, b, v, pol_stream);
                        n_buckets++;
                        bool hole = false;
                        unsigned mbits = 0;
#pragma unroll
                        for (int q = 0; q < 4; q++) { // branch-free classification of the four slots
                            const bool empty = v[q] == kEmptySlot;
                            hole |= empty;
                            // not this read itself: OverlapGraph.cpp:421 / :655
                            const bool match = !empty && (uint32_t)(v[q] >> 33) == tag && ((uint32_t)v[q] >> 1) != (uint32_t)r1;
                            mbits |= (unsigned)match << q;
                        }
                        if (MODE == MODE_EDGES && mbits && p.skip_contained) {
#pragma unroll
                            for (int q = 0; q < 4; q++)
                                if (((mbits >> q) & 1) && is_contained(p.contained_bits, (uint32_t)v[q] >> 1)) mbits &= ~(1u << q);
                        }
                        if (mbits) {
                            const int cnt = __popc(mbits);
                            int pos = atomicAdd(&s.ctrl[0], cnt);
                            pushed += cnt;
#pragma unroll
                            for (int q = 0; q < 4; q++) {
                                if ((mbits >> q) & 1) {
                                    const uint32_t rec = (uint32_t)v[q];
                                    const uint64_t c = make_hit(j, rec, cand_type(rec & 1, (int)((v[q] >> 32) & 1) == fq));
                                    if (pos < hcap) {
                                        s.hits[pos] = c;
                                        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.reads.words + (uint64_t)(rec >> 1) * (uint64_t)p.reads.stride));
                                    } else if (MODE == MODE_CONTAIN) { n_verified++; n_hits += contain_one<NW>(p, s, r1, L1, c); } // queue full (rare)
                                    pos++;
                                }
                            }
                        }
                        if (hole) break;
                        b = (b + 1 == nbuckets) ? 0 : b + 1;
                    }
                    if (MODE == MODE_EDGES && pushed > cap) s.ctrl[3] = 1;
                }
                __syncwarp();
                if (MODE == MODE_CONTAIN) { // ---- 2. verify, drained after every round (this pass has no exact path)
                    const int nc = min(s.ctrl[0], hcap);
                    for (int k = lane; k < nc; k += 32) { n_verified++; n_hits += contain_one<NW>(p, s, r1, L1, s.hits[k]); }
                    __syncwarp();
                    if (lane == 0) s.ctrl[0] = 0;
                    __syncwarp();
                }
            }
            if (MODE != MODE_EDGES) continue;
            if (p.dbg & 4) continue;

            const int nc = s.ctrl[0];
            bool slow = s.ctrl[1] != 0 || nc > hcap;
            int nrow = 0;
            if (!slow) {
                // ---- 2. verify: lane per candidate ------------------------------------------------------------
                const bool count_pos = s.ctrl[3] != 0; // only reads with a crowded position pay for per-position counts
                int hsz = 64;
                while (hsz < 2 * nc) hsz <<= 1;         // <= p.hset
                const uint32_t hm = (uint32_t)hsz - 1;
                for (int k = lane; k < hsz; k += 32) hset[k] = 0xFFFFFFFFu;
                if (count_pos) for (int k = lane; k < jhi; k += 32) cntj[k] = 0;
                __syncwarp();
                bool dup = false, over = false;
                for (int i0 = 0; i0 < nc; i0 += 32) {
                    const int i = i0 + lane;
                    if (i < nc) {
                        const uint64_t c = s.hits[i];
                        const uint32_t r2 = hit_read(c);
                        const bool ok = verify_dovetail<NW>(p, s, L1, hit_j(c), hit_type(c), r2);
                        if (ok) {
                            // ---- 3a. first hit per neighbour: insert r2 into the warp's id set
                            uint32_t hh = (r2 * 0x9E3779B1u) >> 7 & hm;
                            for (;;) {
                                const uint32_t old = atomicCAS(&hset[hh], 0xFFFFFFFFu, r2);
                                if (old == 0xFFFFFFFFu) break;
                                if (old == r2) { dup = true; break; }
                                hh = (hh + 1) & hm;
                            }
                            if (count_pos) over |= atomicAdd(&cntj[hit_j(c)], 1) >= cap;
                        } else {
                            s.hits[i] = ~0ULL;
                        }
                    }
                }
                n_verified += (lane == 0) ? (unsigned)nc : 0u;
                slow = __any_sync(FULL, over); // a position with more than cap partners: redo exactly
                dup = __any_sync(FULL, dup);
                __syncwarp();
                if (!slow && dup) {
                    // rare (tandem repeats, circular overlaps): a neighbour reached through two positions keeps the
                    // first one in (position, record) order (OverlapGraph.cpp:656)
                    unsigned drop = 0;
                    for (int i0 = 0, rd = 0; i0 < nc; i0 += 32, rd++) {
                        const int i = i0 + lane;
                        const uint64_t hk = (i < nc) ? s.hits[i] : ~0ULL;
                        if (hk == ~0ULL) continue;
                        const uint32_t r2 = hit_read(hk);
                        for (int k = 0; k < nc; k++) {
                            const uint64_t o = s.hits[k];
                            if (o != ~0ULL && hit_read(o) == r2 && o < hk) { drop |= 1u << rd; break; }
                        }
                    }
                    __syncwarp();
                    for (int i0 = 0, rd = 0; i0 < nc; i0 += 32, rd++)
                        if ((drop >> rd) & 1) s.hits[i0 + lane] = ~0ULL;
                    __syncwarp();
                }
                if (!slow) {
                    for (int i0 = 0; i0 < nc; i0 += 32) {
                        const int i = i0 + lane;
                        nrow += __popc(__ballot_sync(FULL, i < nc && s.hits[i] != ~0ULL));
                    }
                    n_hits += (lane == 0) ? (unsigned)nrow : 0u;
                }
            }
            if (slow) {
                n_slow += (lane == 0);
                search_edges_slow<NW>(p, s, r1, L1, lane, n_probes, n_buckets, n_verified, n_capfired);
                nrow = s.ctrl[2];
                if (nrow > p.rowcap) nrow = p.rowcap; // cannot happen: rowcap >= cap * positions
            }
            if (nrow == 0) continue;
            // ---- 3b. append the row to the global adjacency.  Space comes from a warp-private slice reserved
            // kRowBlock entries at a time: one global atomic per ~30 reads instead of one per read (rows need not be
            // contiguous in read order; rowinfo points at them).  Rows are stored unsorted: the reduction picks
            // neighbours in offset order itself.
            if (blk_cur + nrow > blk_end) {
                const unsigned long long want = nrow > kRowBlock ? (unsigned long long)nrow : (unsigned long long)kRowBlock;
                if (lane == 0) blk_cur = atomicAdd(p.rows_cursor, want);
                blk_cur = __shfl_sync(FULL, blk_cur, 0);
                blk_end = blk_cur + want;
            }
            const unsigned long long base = blk_cur;
            blk_cur += nrow;
            n_entries += (lane == 0) ? (unsigned long long)nrow : 0ULL;
            if ((unsigned)nrow > maxdeg) maxdeg = nrow;
            if (base + nrow <= p.rows_cap) {
                if (slow) {
                    for (int i = lane; i < nrow; i += 32) p.rows[base + i] = s.row[i];
                } else {
                    int off = 0;
                    for (int i0 = 0; i0 < nc; i0 += 32) {
                        const int i = i0 + lane;
                        const uint64_t hk = (i < nc) ? s.hits[i] : ~0ULL;
                        const bool valid = hk != ~0ULL;
                        const unsigned m = __ballot_sync(FULL, valid);
                        if (valid) {
                            int orient, ovl;
                            type_to_edge(hit_type(hk), L1, K, hit_j(hk), &orient, &ovl);
                            p.rows[base + off + __popc(m & lt_mask)] = make_entry(L1 - ovl, hit_read(hk), orient);
                        }
                        off += __popc(m);
                    }
                }
                if (lane == 0) p.rowinfo[r1] = make_rowinfo(base, (uint32_t)nrow);
            } else if (lane == 0) {
                atomicExch(p.stats + ST_OVERFLOW, 1ULL);
            }
            __syncwarp();
        }
    }
    warp_stat_add(p.stats, ST_QUERIES, n_queries);
    warp_stat_add(p.stats, ST_PROBES, n_probes);
    warp_stat_add(p.stats, ST_BUCKETS, n_buckets);
    warp_stat_add(p.stats, ST_VERIFIED, n_verified);
    warp_stat_add(p.stats, ST_HITS, n_hits);
    if (MODE == MODE_EDGES) {
        warp_stat_add(p.stats, ST_ENTRIES, n_entries);
        warp_stat_add(p.stats, ST_CAP_FIRED, n_capfired);
        warp_stat_add(p.stats, ST_SLOW_READS, n_slow);
        for (int o = 16; o; o >>= 1) { unsigned t = __shfl_xor_sync(FULL, maxdeg, o); if (t > maxdeg) maxdeg = t; }
        if (lane == 0 && maxdeg) atomicMax(p.stats + ST_MAXDEG, (unsigned long long)maxdeg);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Edge pass as three kernels (default).  The fused kernel above needs 64+ registers and 5 KB of shared memory per
// warp, which caps it at 32 warps per SM while most of its time is memory latency; split, each part is small:
//   k_edges_probe  : stage, hash, presence filter, bucket probe  -> tag matches parked in the read's (provisional)
//                    adjacency row in global memory                                   (40 registers, 48 warps / SM)
//   k_edges_verify : stage, load the parked candidates, lane-per-candidate overlap compare, first-hit-per-neighbour,
//                    compact the survivors in place                                   (no table code)
//   k_edges_exact  : the few reads flagged for the exact sequential search (cap may fire, long chains, > hcap
//                    candidates)
// ---------------------------------------------------------------------------------------------------------------
__host__ __device__ inline size_t probe_words_per_warp(int WP, int npos, int hcap) { return 2 * (size_t)WP + (size_t)npos + ((size_t)npos + 1) / 2 + (size_t)hcap + 4; }
__host__ __device__ inline size_t verify_words_per_group(int WP, int npos, int hcap, int hset, int gw)
{   // A, R, candidate queue, id set + per-position counters, transposed candidate words (gw x up to 16), control
    const int nw = WP - 2 <= 16 ? ((WP - 2 + 1) / 2) * 2 : 0;
    return 2 * (size_t)WP + (size_t)hcap + ((size_t)hset + (size_t)npos + 1) / 2 + (size_t)gw * nw + 2;
}
__host__ __device__ inline size_t exact_words_per_warp(int WP, int rowcap) { return 2 * (size_t)WP + (size_t)rowcap + kBestMax + 2; }

// SHARDED: the table is key-sharded over GPUs, buckets are read through NVLink peer pointers (a separate instantiation
// so that the single-GPU kernel keeps its registers)
template <bool SHARDED>
__global__ void __launch_bounds__(kThreads, 6) k_edges_probe(SearchParams p)
{
    extern __shared__ uint64_t smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int WP = ((p.reads.max_len + 31) >> 5) + 2;
    const int K = p.K;
    uint64_t *w0 = smem + wib * probe_words_per_warp(WP, p.npos, p.hcap);
    uint32_t *A = reinterpret_cast<uint32_t *>(w0), *R = A + 2 * WP;
    uint64_t *ph = w0 + 2 * WP;
    uint32_t *pj = reinterpret_cast<uint32_t *>(ph + p.npos);
    uint64_t *hits = ph + p.npos + (p.npos + 1) / 2;
    int *ctrl = reinterpret_cast<int *>(hits + p.hcap);
    const unsigned lt_mask = (1u << lane) - 1;
    const uint64_t nbuckets = p.table.nbuckets;
    const int hcap = p.hcap, cap = p.cap;
    unsigned n_queries = 0, n_probes = 0, n_buckets = 0;
    unsigned long long blk_cur = 0, blk_end = 0; // this warp's reserved slice of the adjacency buffer
    const bool use_pre = p.reads.stride <= 32;
    uint64_t pre_for = ~0ULL, pre_word = 0;
    const uint64_t pol_stream = policy_evict_first(), pol_keep = policy_evict_last();
    uint64_t rb, re;
    while (grab_chunk(p.work_counter, p.q_lo, p.q_hi, lane, &rb, &re)) {
        for (uint64_t r1 = rb; r1 < re; r1++) {
            if ((__ldg(p.contained_bits + (r1 >> 5)) >> (r1 & 31)) & 1) continue; // OverlapGraph.cpp:657
            const int L1 = read_len(p.reads, r1);
            if (use_pre) {
                const uint64_t mine = (pre_for == r1) ? pre_word : ((lane < p.reads.stride) ? __ldg(p.reads.words + r1 * (uint64_t)p.reads.stride + lane) : 0ULL);
                if (r1 + 1 < re) { pre_for = r1 + 1; pre_word = (lane < p.reads.stride) ? __ldg(p.reads.words + (r1 + 1) * (uint64_t)p.reads.stride + lane) : 0ULL; }
                stage_read_pre(mine, L1, A, R, WP, lane);
            } else {
                stage_read(p.reads, r1, L1, A, R, WP, lane);
            }
            n_queries += (lane == 0);
            if (lane < 8) ctrl[lane] = 0; // [0] queued, [1] needs exact path, [3] some position has > cap candidates, [4..7] parking
            __syncwarp();
            const int jhi = L1 - K; // positions [1, L1-K) (OverlapGraph.cpp:638)
            // ---- hash + presence filter: passing positions ballot-compacted, their bucket prefetched into L2
            int np = 0;
            for (int jb = 1; jb < jhi; jb += 32) {
                const int j = jb + lane;
                uint64_t h = 0;
                int fq = 0;
                bool pass = false;
                if (j < jhi) {
                    h = canon_kmer_hash(A, R, L1, j, K, &fq);
                    n_probes++;
                    pass = filter_test(p.table, h, pol_keep);
                    if (!SHARDED && pass) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.table.slots + 4 * bucket_of(h, nbuckets)));
                }
                const unsigned m = __ballot_sync(FULL, pass);
                if (pass) {
                    const int pos = np + __popc(m & lt_mask);
                    ph[pos] = h;
                    pj[pos] = ((uint32_t)j << 1) | (uint32_t)fq;
                }
                np += __popc(m);
            }
            __syncwarp();
            // ---- probe: one lane per surviving position.  Tag matches go to the shared-memory queue; a read with many
            // of them (high coverage) flushes the queue into a kParkMax-entry row reserved on first need.
            // (the parking state lives in shared memory -- ctrl[4] parked, ctrl[5] capacity, ctrl[6..7] row start -- so that
            // the common case, no flush at all, keeps the registers for the probe loop)
            bool overflow = false;
            auto reserve = [&](int n) -> unsigned long long { // n entries from this warp's slice of the adjacency buffer
                if (blk_cur + n > blk_end) {
                    const unsigned long long want = n > kRowBlock ? (unsigned long long)n : (unsigned long long)kRowBlock;
                    if (lane == 0) blk_cur = atomicAdd(p.rows_cursor, want);
                    blk_cur = __shfl_sync(FULL, blk_cur, 0);
                    blk_end = blk_cur + want;
                }
                const unsigned long long b = blk_cur;
                blk_cur += n;
                return b;
            };
            for (int i0 = 0; i0 < np; i0 += 32) {
                const int i = i0 + lane;
                if (i < np) {
                    const uint64_t h = ph[i];
                    const uint32_t jf = pj[i];
                    const int j = (int)(jf >> 1), fq = (int)(jf & 1);
                    const uint32_t tag = slot_tag(h);
                    const uint64_t *slots = p.table.slots;
                    uint64_t b;
                    if (SHARDED) { const Home home = home_of(p.table, h); slots = home.slots; b = home.b; }
                    else b = bucket_of(h, nbuckets);
                    int pushed = 0;
                    for (int walked = 0;; walked++) {
                        if (walked == kScanLimit) { ctrl[1] = 1; break; } // long chain: exact path
                        uint64_t v[4];
                        load_bucket(slots, b, v, pol_stream);
                        n_buckets++;
                        bool hole = false;
                        unsigned mbits = 0;
#pragma unroll
                        for (int q = 0; q < 4; q++) { // branch-free classification of the four slots
                            const bool empty = v[q] == kEmptySlot;
                            hole |= empty;
                            const bool match = !empty && (uint32_t)(v[q] >> 33) == tag && ((uint32_t)v[q] >> 1) != (uint32_t)r1; // :655
                            mbits |= (unsigned)match << q;
                        }
                        if (mbits && p.skip_contained) { // "ignore contained reads" (HashTable.cpp:533); the bitmap sits in L2
#pragma unroll
                            for (int q = 0; q < 4; q++)
                                if (((mbits >> q) & 1) && is_contained(p.contained_bits, (uint32_t)v[q] >> 1)) mbits &= ~(1u << q);
                        }
                        if (mbits) {
                            const int cnt = __popc(mbits);
                            int pos = atomicAdd(&ctrl[0], cnt);
                            pushed += cnt;
#pragma unroll
                            for (int q = 0; q < 4; q++) {
                                if ((mbits >> q) & 1) {
                                    const uint32_t rec = (uint32_t)v[q];
                                    if (pos < hcap) hits[pos] = make_hit(j, rec, cand_type(rec & 1, (int)((v[q] >> 32) & 1) == fq));
                                    pos++;
                                }
                            }
                        }
                        if (hole) break;
                        b = (b + 1 == nbuckets) ? 0 : b + 1;
                    }
                    if (pushed > cap) ctrl[3] = 1;
                }
                __syncwarp();
                const int q = ctrl[0];
                if (q > hcap) { overflow = true; break; }            // one round overran the queue: exact path
                if (q >= hcap / 2 && i0 + 32 < np) {                 // make room for the next round
                    unsigned long long *pb = reinterpret_cast<unsigned long long *>(ctrl + 6);
                    if (!ctrl[5]) {
                        const unsigned long long b = reserve(kParkMax);
                        __syncwarp();
                        if (lane == 0) { *pb = b; ctrl[5] = kParkMax; }
                        __syncwarp();
                    }
                    const int parked = ctrl[4];
                    if (parked + q > ctrl[5]) { overflow = true; break; }
                    if (*pb + ctrl[5] <= p.rows_cap)
                        for (int k = lane; k < q; k += 32) p.rows[*pb + parked + k] = hits[k];
                    __syncwarp();
                    if (lane == 0) { ctrl[0] = 0; ctrl[4] = parked + q; }
                    __syncwarp();
                }
            }
            __syncwarp();
            // ---- park the (remaining) candidates in the read's row (the verify kernel compacts the survivors in place)
            const int q = ctrl[0];
            const int parked = ctrl[4], park_cap = ctrl[5];
            const int nc = parked + q;
            if (overflow || ctrl[1] != 0 || q > hcap || (park_cap && nc > park_cap)) {
                if (lane == 0) p.rowinfo[r1] = kInfoExact;
            } else if (nc > 0) {
                const unsigned long long base = park_cap ? *reinterpret_cast<unsigned long long *>(ctrl + 6) : reserve(nc);
                const unsigned long long room = park_cap ? (unsigned long long)park_cap : (unsigned long long)nc;
                if (base + room <= p.rows_cap) {
                    for (int k = lane; k < q; k += 32) p.rows[base + parked + k] = hits[k];
                    if (lane == 0) p.rowinfo[r1] = make_rowinfo(base, (uint32_t)nc) | (ctrl[3] ? kInfoCrowded : 0ULL);
                } else if (lane == 0) {
                    atomicExch(p.stats + ST_OVERFLOW, 1ULL);
                }
            }
            __syncwarp();
        }
    }
    warp_stat_add(p.stats, ST_QUERIES, n_queries);
    warp_stat_add(p.stats, ST_PROBES, n_probes);
    warp_stat_add(p.stats, ST_BUCKETS, n_buckets);
}

// GW = lanes that work on one read (32, or 16: two reads per warp side by side -- a read has ~36 candidates, so 32-lane
// rounds leave the second one almost empty; 16-lane groups keep ~80% of the lanes busy and halve the per-read overhead)
template <int NW, int GW>
__global__ void __launch_bounds__(kThreads, (NW > 0 && NW <= 8) ? (GW == 32 ? 5 : 4) : 1) k_edges_verify(SearchParams p)
{
    extern __shared__ uint64_t smem[];
    const int lane = threadIdx.x & (GW - 1), gib = threadIdx.x / GW;   // lane in the group, group in the block
    const int gshift = (threadIdx.x & 31) & ~(GW - 1);                 // the group's first lane within the warp
    const unsigned gmask = GW == 32 ? FULL : (((1u << GW) - 1u) << gshift);
    const int WP = ((p.reads.max_len + 31) >> 5) + 2;
    const int K = p.K;
    WarpSmem s;
    uint64_t *w0 = smem + gib * verify_words_per_group(WP, p.npos, p.hcap, p.hset, GW);
    s.A = reinterpret_cast<uint32_t *>(w0); s.R = s.A + 2 * WP; s.p_u32 = 2 * WP;
    s.hits = w0 + 2 * WP;
    uint32_t *hset = reinterpret_cast<uint32_t *>(s.hits + p.hcap);
    int *cntj = reinterpret_cast<int *>(hset + p.hset);
    s.ctrl = reinterpret_cast<int *>(w0 + verify_words_per_group(WP, p.npos, p.hcap, p.hset, GW) - 2);
    uint64_t *cw = w0 + 2 * WP + p.hcap + (p.hset + p.npos + 1) / 2 + lane; // transposed candidate words, this lane's column
    s.row = nullptr; s.best = nullptr; s.ph = nullptr; s.pj = nullptr;
    const unsigned lt_mask = (1u << lane) - 1;
    const int cap = p.cap;
    unsigned n_verified = 0, n_hits = 0, maxdeg = 0;
    unsigned long long n_entries = 0;
    uint64_t rb, re;
    while (grab_chunk_g<GW>(p.work_counter + 1, p.q_lo, p.q_hi, lane, gmask, &rb, &re)) {
        for (uint64_t r1 = rb; r1 < re; r1++) {
            const uint64_t ri = p.rowinfo[r1];
            const int nc = (int)rowinfo_deg(ri);
            if ((ri & kInfoExact) || nc == 0) continue;
            const uint64_t start = rowinfo_start(ri);
            const int L1 = read_len(p.reads, r1);
            if (nc > p.hcap) {
                // ---- big row (high coverage, up to kParkMax candidates): same steps, but the candidates stay in the
                // parked row in global memory and are taken GW at a time
                stage_read_g<GW>(p.reads, r1, L1, s.A, s.R, WP, lane, gmask);
                const bool count_pos = (ri & kInfoCrowded) != 0;
                if (count_pos) for (int k = lane; k < L1 - K; k += GW) cntj[k] = 0;
                __syncwarp(gmask);
                bool over = false;
                for (int i0 = 0; i0 < nc; i0 += GW) {
                    const int i = i0 + lane;
                    if (i < nc) {
                        const uint64_t c = p.rows[start + i];
                        const bool ok = verify_dovetail<NW>(p, s, L1, hit_j(c), hit_type(c), hit_read(c));
                        if (!ok) p.rows[start + i] = ~0ULL;
                        else if (count_pos) over |= atomicAdd(&cntj[hit_j(c)], 1) >= cap;
                    }
                }
                n_verified += (lane == 0) ? (unsigned)nc : 0u;
                if (__any_sync(gmask, over)) { // a position with more than cap partners: redo exactly
                    if (lane == 0) p.rowinfo[r1] = kInfoExact;
                    __syncwarp(gmask);
                    continue;
                }
                __syncwarp(gmask);
                // first hit per neighbour (OverlapGraph.cpp:656): neighbour ids go through a shared-memory set in two
                // passes (one hash bit each) so that 512 slots are enough; an id seen twice marks the read for the
                // quadratic clean-up below
                uint32_t *bigset = reinterpret_cast<uint32_t *>(s.hits); // queue + id-set space, unused on this path
                const int slots = 2 * p.hcap + p.hset;                    // u32 slots (hits is u64[hcap])
                bool dup = false, full = false;
                for (int pass = 0; pass < 2; pass++) {
                    for (int k = lane; k < slots; k += GW) bigset[k] = 0xFFFFFFFFu;
                    __syncwarp(gmask);
                    for (int i0 = 0; i0 < nc; i0 += GW) {
                        const int i = i0 + lane;
                        const uint64_t hk = (i < nc) ? __ldcg(p.rows + start + i) : ~0ULL;
                        if (hk == ~0ULL) continue;
                        const uint32_t r2 = hit_read(hk);
                        const uint32_t hv = r2 * 0x9E3779B1u;
                        if ((int)(hv >> 31) != pass) continue;
                        uint32_t hh = (hv >> 7) % (uint32_t)slots;
                        for (int tries = 0;; tries++) {
                            if (tries == slots) { full = true; break; }
                            const uint32_t old = atomicCAS(&bigset[hh], 0xFFFFFFFFu, r2);
                            if (old == 0xFFFFFFFFu) break;
                            if (old == r2) { dup = true; break; }
                            hh = (hh + 1 == (uint32_t)slots) ? 0 : hh + 1;
                        }
                    }
                    __syncwarp(gmask);
                }
                if (__any_sync(gmask, dup || full)) {
                    unsigned drop = 0;
                    for (int i0 = 0, rd = 0; i0 < nc; i0 += GW, rd++) {
                        const int i = i0 + lane;
                        const uint64_t hk = (i < nc) ? __ldcg(p.rows + start + i) : ~0ULL;
                        if (hk == ~0ULL) continue;
                        const uint32_t r2 = hit_read(hk);
                        for (int k = 0; k < nc; k++) {
                            const uint64_t o = __ldcg(p.rows + start + k);
                            if (o != ~0ULL && hit_read(o) == r2 && o < hk) { drop |= 1u << rd; break; }
                        }
                    }
                    __syncwarp(gmask);
                    for (int i0 = 0, rd = 0; i0 < nc; i0 += GW, rd++)
                        if ((drop >> rd) & 1) p.rows[start + i0 + lane] = ~0ULL;
                    __syncwarp(gmask);
                }
                // compact the survivors at the front of the row, as adjacency entries
                int off = 0;
                for (int i0 = 0; i0 < nc; i0 += GW) {
                    const int i = i0 + lane;
                    const uint64_t hk = (i < nc) ? __ldcg(p.rows + start + i) : ~0ULL;
                    const bool valid = hk != ~0ULL;
                    const unsigned m = (__ballot_sync(gmask, valid) >> gshift); // everybody has read before anybody writes
                    if (valid) {
                        int orient, ovl;
                        type_to_edge(hit_type(hk), L1, K, hit_j(hk), &orient, &ovl);
                        p.rows[start + off + __popc(m & lt_mask)] = make_entry(L1 - ovl, hit_read(hk), orient);
                    }
                    off += __popc(m);
                    __syncwarp(gmask);
                }
                if (lane == 0) p.rowinfo[r1] = off ? make_rowinfo(start, (uint32_t)off) : 0ULL;
                n_hits += (lane == 0) ? (unsigned)off : 0u;
                n_entries += (lane == 0) ? (unsigned long long)off : 0ULL;
                if ((unsigned)off > maxdeg) maxdeg = off;
                __syncwarp(gmask);
                continue;
            }
            // candidates first (their rows are the long-latency loads of this kernel), then the query
            // (split: a candidate whose suffix overlaps is read from the reverse-complement copy, where that suffix is
            // the prefix -- so only the leading sector(s) the overlap reaches are touched at all)
            const bool split = NW >= 4 && p.reads.words_rc != nullptr;
            for (int i = lane; i < nc; i += GW) {
                const uint64_t c = p.rows[start + i];
                s.hits[i] = c;
                const int t = hit_type(c);
                const uint64_t *row = ((split && (t == 1 || t == 2)) ? p.reads.words_rc : p.reads.words) + (uint64_t)hit_read(c) * (uint64_t)p.reads.stride;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(row));
                if (split && NW > 4 && ((t == 0 || t == 2) ? L1 - hit_j(c) : K + hit_j(c)) > 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(row + 4));
            }
            stage_read_g<GW>(p.reads, r1, L1, s.A, s.R, WP, lane, gmask);
            const bool count_pos = (ri & kInfoCrowded) != 0; // only reads with a crowded position pay for per-position counts
            int hsz = 64;
            while (hsz < 2 * nc) hsz <<= 1;         // <= p.hset
            const uint32_t hm = (uint32_t)hsz - 1;
            for (int k = lane; k < hsz; k += GW) hset[k] = 0xFFFFFFFFu;
            if (count_pos) for (int k = lane; k < L1 - K; k += GW) cntj[k] = 0;
            __syncwarp(gmask);
            bool dup = false, over = false;
            for (int i0 = 0; i0 < nc; i0 += GW) {
                const int i = i0 + lane;
                if (i < nc) {
                    const uint64_t c = s.hits[i];
                    const uint32_t r2 = hit_read(c);
                    bool ok;
                    if (NW == 0) {
                        ok = verify_dovetail<NW>(p, s, L1, hit_j(c), hit_type(c), r2);
                    } else if (split) {
                        const int L2 = read_len(p.reads, r2), t = hit_type(c);
                        int use_rc, a, b, n;
                        ok = dovetail_window(t, L1, hit_j(c), K, L2, &use_rc, &a, &b, &n);
                        if (ok) {
                            const uint64_t *row = p.reads.words;
                            // s2[L2-n..L2) == X  <=>  rc(s2)[0..n) == rc(X): swap the query array, mirror the window
                            if (t == 1 || t == 2) { row = p.reads.words_rc; use_rc ^= 1; a = L1 - a - n; }
                            row += (uint64_t)r2 * (uint64_t)p.reads.stride;
                            const uint64_t pol = policy_evict_first();
#pragma unroll
                            for (int sct = 0; sct < ((NW > 0 ? NW : 4) + 3) / 4; sct++) {
                                if (sct * 128 < n) { // one 32-byte sector = 128 bases, one LDG.E.256
                                    uint64_t v[4];
                                    asm volatile("ld.global.nc.L2::cache_hint.v4.u64 {%0,%1,%2,%3}, [%4], %5;"
                                                 : "=l"(v[0]), "=l"(v[1]), "=l"(v[2]), "=l"(v[3]) : "l"(row + 4 * sct), "l"(pol));
#pragma unroll
                                    for (int k = 0; k < 4; k++)
                                        if (4 * sct + k < NW) cw[(4 * sct + k) * GW] = v[k];
                                }
                            }
                            ok = SmemMatcher<GW>{cw}(use_rc ? s.R : s.A, a, 0, n);
                        }
                    } else {
                        RegMatcher<(NW > 0 ? NW : 2)> m;
                        m.stride = p.reads.stride;
                        m.load(p.reads.words, r2);
                        const int L2 = read_len(p.reads, r2);
#pragma unroll
                        for (int w = 0; w < (NW > 0 ? NW : 2); w++) cw[w * GW] = m.v[w];
                        int use_rc, a, b, n;
                        ok = dovetail_window(hit_type(c), L1, hit_j(c), K, L2, &use_rc, &a, &b, &n) &&
                             SmemMatcher<GW>{cw}(use_rc ? s.R : s.A, a, b, n);
                    }
                    if (ok) {
                        // first hit per neighbour: insert r2 into the warp's id set
                        uint32_t hh = (r2 * 0x9E3779B1u) >> 7 & hm;
                        for (;;) {
                            const uint32_t old = atomicCAS(&hset[hh], 0xFFFFFFFFu, r2);
                            if (old == 0xFFFFFFFFu) break;
                            if (old == r2) { dup = true; break; }
                            hh = (hh + 1) & hm;
                        }
                        if (count_pos) over |= atomicAdd(&cntj[hit_j(c)], 1) >= cap;
                    } else {
                        s.hits[i] = ~0ULL;
                    }
                }
            }
            n_verified += (lane == 0) ? (unsigned)nc : 0u;
            over = __any_sync(gmask, over); // a position with more than cap partners: redo exactly
            dup = __any_sync(gmask, dup);
            __syncwarp(gmask);
            if (over) {
                if (lane == 0) p.rowinfo[r1] = kInfoExact;
                continue;
            }
            if (dup) {
                // rare (tandem repeats, circular overlaps): a neighbour reached through two positions keeps the
                // first one in (position, record) order (OverlapGraph.cpp:656)
                unsigned drop = 0;
                for (int i0 = 0, rd = 0; i0 < nc; i0 += GW, rd++) {
                    const int i = i0 + lane;
                    const uint64_t hk = (i < nc) ? s.hits[i] : ~0ULL;
                    if (hk == ~0ULL) continue;
                    const uint32_t r2 = hit_read(hk);
                    for (int k = 0; k < nc; k++) {
                        const uint64_t o = s.hits[k];
                        if (o != ~0ULL && hit_read(o) == r2 && o < hk) { drop |= 1u << rd; break; }
                    }
                }
                __syncwarp(gmask);
                for (int i0 = 0, rd = 0; i0 < nc; i0 += GW, rd++)
                    if ((drop >> rd) & 1) s.hits[i0 + lane] = ~0ULL;
                __syncwarp(gmask);
            }
            // survivors become adjacency entries, ballot-compacted at the front of the same row (unsorted: the
            // reduction picks neighbours in offset order itself)
            int off = 0;
            for (int i0 = 0; i0 < nc; i0 += GW) {
                const int i = i0 + lane;
                const uint64_t hk = (i < nc) ? s.hits[i] : ~0ULL;
                const bool valid = hk != ~0ULL;
                const unsigned m = (__ballot_sync(gmask, valid) >> gshift);
                if (valid) {
                    int orient, ovl;
                    type_to_edge(hit_type(hk), L1, K, hit_j(hk), &orient, &ovl);
                    p.rows[start + off + __popc(m & lt_mask)] = make_entry(L1 - ovl, hit_read(hk), orient);
                }
                off += __popc(m);
            }
            if (lane == 0) p.rowinfo[r1] = off ? make_rowinfo(start, (uint32_t)off) : 0ULL;
            n_hits += (lane == 0) ? (unsigned)off : 0u;
            n_entries += (lane == 0) ? (unsigned long long)off : 0ULL;
            if ((unsigned)off > maxdeg) maxdeg = off;
            __syncwarp(gmask);
        }
    }
    warp_stat_add(p.stats, ST_VERIFIED, n_verified);
    warp_stat_add(p.stats, ST_HITS, n_hits);
    warp_stat_add(p.stats, ST_ENTRIES, n_entries);
    for (int o = 16; o; o >>= 1) { unsigned t = __shfl_xor_sync(FULL, maxdeg, o); if (t > maxdeg) maxdeg = t; }
    if (lane == 0 && maxdeg) atomicMax(p.stats + ST_MAXDEG, (unsigned long long)maxdeg);
}

template <int NW>
__global__ void __launch_bounds__(kThreads) k_edges_exact(SearchParams p)
{
    extern __shared__ uint64_t smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int WP = ((p.reads.max_len + 31) >> 5) + 2;
    WarpSmem s;
    uint64_t *w0 = smem + wib * exact_words_per_warp(WP, p.rowcap);
    s.A = reinterpret_cast<uint32_t *>(w0); s.R = s.A + 2 * WP; s.p_u32 = 2 * WP;
    s.row = w0 + 2 * WP; s.best = s.row + p.rowcap; s.ctrl = reinterpret_cast<int *>(s.best + kBestMax);
    s.hits = nullptr; s.ph = nullptr; s.pj = nullptr;
    unsigned n_probes = 0, n_buckets = 0, n_verified = 0, n_capfired = 0, n_slow = 0, maxdeg = 0;
    unsigned long long n_entries = 0;
    uint64_t rb, re;
    while (grab_chunk(p.work_counter + 2, p.q_lo, p.q_hi, lane, &rb, &re)) {
        for (uint64_t r1 = rb; r1 < re; r1++) {
            if (!(p.rowinfo[r1] & kInfoExact)) continue;
            const int L1 = read_len(p.reads, r1);
            stage_read(p.reads, r1, L1, s.A, s.R, WP, lane);
            n_slow += (lane == 0);
            search_edges_slow<NW>(p, s, r1, L1, lane, n_probes, n_buckets, n_verified, n_capfired);
            int nrow = s.ctrl[2];
            if (nrow > p.rowcap) nrow = p.rowcap; // cannot happen: rowcap >= cap * positions
            unsigned long long base = 0;
            if (lane == 0 && nrow) base = atomicAdd(p.rows_cursor, (unsigned long long)nrow);
            base = __shfl_sync(FULL, base, 0);
            if (nrow == 0) { if (lane == 0) p.rowinfo[r1] = 0ULL; continue; }
            n_entries += (lane == 0) ? (unsigned long long)nrow : 0ULL;
            if ((unsigned)nrow > maxdeg) maxdeg = nrow;
            if (base + nrow <= p.rows_cap) {
                for (int i = lane; i < nrow; i += 32) p.rows[base + i] = s.row[i];
                if (lane == 0) p.rowinfo[r1] = make_rowinfo(base, (uint32_t)nrow);
            } else if (lane == 0) {
                p.rowinfo[r1] = 0ULL;
                atomicExch(p.stats + ST_OVERFLOW, 1ULL);
            }
            __syncwarp();
        }
    }
    warp_stat_add(p.stats, ST_PROBES, n_probes);
    warp_stat_add(p.stats, ST_BUCKETS, n_buckets);
    warp_stat_add(p.stats, ST_VERIFIED, n_verified);
    warp_stat_add(p.stats, ST_ENTRIES, n_entries);
    warp_stat_add(p.stats, ST_CAP_FIRED, n_capfired);
    warp_stat_add(p.stats, ST_SLOW_READS, n_slow);
    for (int o = 16; o; o >>= 1) { unsigned t = __shfl_xor_sync(FULL, maxdeg, o); if (t > maxdeg) maxdeg = t; }
    if (lane == 0 && maxdeg) atomicMax(p.stats + ST_MAXDEG, (unsigned long long)maxdeg);
}

} // namespace disco
#include "edges_flat.cuh"
namespace disco {

// ---------------------------------------------------------------------------------------------------------------
// containment bookkeeping
// ---------------------------------------------------------------------------------------------------------------
// One reservation per BLOCK in a global cursor for the items its warps flagged (m = the warp's ballot): a single hot
// counter takes one atomic per 256 reads instead of one per warp (same-address atomics serialise: with a tenth of the reads
// contained nearly every warp has one, and at 80 M reads the three bookkeeping kernels spent 4 ms on their counters).
// Every thread of the block must call it; returns where this warp's items start.
__device__ __forceinline__ unsigned long long block_reserve(unsigned m, unsigned long long *cursor)
{
    __shared__ unsigned wcnt[32];
    __shared__ unsigned long long sbase;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    if (lane == 0) wcnt[wib] = (unsigned)__popc(m);
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned tot = 0;
        for (int w = 0; w < nw; w++) { const unsigned t = wcnt[w]; wcnt[w] = tot; tot += t; }
        sbase = tot ? atomicAdd(cursor, (unsigned long long)tot) : 0ULL;
    }
    __syncthreads();
    const unsigned long long at = sbase + wcnt[wib];
    __syncthreads(); // (the arrays are reused by the caller's next reservation)
    return at;
}

__global__ void k_contained_finish(const unsigned long long *best, uint64_t n, uint32_t *bits, unsigned long long *count)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool c = i < n && best[i] != ~0ULL;
    const unsigned m = __ballot_sync(FULL, c);
    if ((threadIdx.x & 31) == 0 && (i >> 5) < ((n + 31) >> 5)) bits[i >> 5] = m;
    block_reserve(m, count);
}

__global__ void k_contained_rows(const unsigned long long *best, ReadsView rv, int K, disco_crow *out, unsigned long long *cursor)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long key = i < rv.n ? best[i] : ~0ULL;
    const bool c = key != ~0ULL;
    const unsigned m = __ballot_sync(FULL, c);
    const int lane = threadIdx.x & 31;
    const unsigned long long base = block_reserve(m, cursor);
    if (c) {
        const uint64_t r1 = key >> 20;
        const int j = (int)((key >> 4) & 0xFFFF), type = (int)(key & 3);
        const int L1 = read_len(rv, r1);
        int orient, ovl;
        type_to_edge(type, L1, K, j, &orient, &ovl);
        disco_crow row;
        row.contained = (uint32_t)i; row.container = (uint32_t)r1; row.orient = (uint32_t)orient; row.start = (uint32_t)(L1 - ovl);
        out[base + __popc(m & ((1u << lane) - 1))] = row;
    }
}

// the contained rows whose read lies in [lo, hi): a rank of a multi-GPU run hands out the rows of its own range only
__global__ void k_crows_in_range(const disco_crow *in, uint64_t n, uint32_t lo, uint32_t hi, disco_crow *out, unsigned long long *cursor)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    disco_crow r{};
    if (i < n) r = in[i];
    const bool c = i < n && r.contained >= lo && r.contained < hi;
    const unsigned m = __ballot_sync(FULL, c);
    const unsigned long long base = block_reserve(m, cursor);
    if (c) out[base + __popc(m & ((1u << (threadIdx.x & 31)) - 1))] = r;
}

// packed reads arrive with the caller's row pitch; the kernels want power-of-two rows (one DRAM line per candidate)
__global__ void k_restride(const uint64_t *src, int src_stride, int src_words, uint64_t *dst, int dst_stride, uint64_t n)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; // one destination word per thread
    const uint64_t r = i / dst_stride;
    const int w = (int)(i - r * dst_stride);
    if (r < n) dst[i] = (w < src_words) ? src[r * src_stride + w] : 0ULL;
}

// reverse complement of every read, same row layout.  A dovetail overlap always covers a prefix or a suffix of the
// candidate; the suffix of a read is the prefix of its reverse complement, so with this copy the verify kernel only ever
// needs the leading 32-byte sector(s) of a row -- one random 32-byte access instead of a 64-byte one for overlaps of up
// to 128 bases (measured random-access rates: 39 G/s at 32 bytes, 22 G/s at 64 bytes).
__global__ void k_revcomp_rows(ReadsView rv, uint64_t *out)
{
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rv.n) return;
    const int L = read_len(rv, r), W = (L + 31) >> 5, S = rv.stride;
    const uint64_t *src = rv.words + r * (uint64_t)S;
    uint64_t *dst = out + r * (uint64_t)S;
    const int sh = (32 * W - L) * 2; // pad bases of the last word, shifted out of the reversed string
    for (int w = 0; w < S; w++) {
        uint64_t v = 0;
        if (w < W) {
            const uint64_t z0 = revcomp64(__ldg(src + (W - 1 - w)));
            const uint64_t z1 = (sh && w + 1 < W) ? revcomp64(__ldg(src + (W - 2 - w))) : 0ULL;
            v = sh ? ((z0 << sh) | (z1 >> (64 - sh))) : z0;
        }
        dst[w] = v;
    }
}

// Tail sector: the 128 bases that end every read, as one 32-byte sector (32 bytes per read).  A dovetail overlap covers a
// prefix or a suffix of the candidate; with this copy a suffix of up to 128 bases is ONE 32-byte access, like a prefix
// is in the row itself (measured random-access rates: 39 G/s at 32 bytes, 22 G/s at 64 bytes).
__global__ void k_make_tails(ReadsView rv, uint64_t *out)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t r = i >> 2;
    const int t = (int)(i & 3);
    if (r >= rv.n) return;
    const int L = read_len(rv, r);
    const int p0 = L - 128 + 32 * t; // first base of this word (negative: in front of the read)
    const uint64_t *src = rv.words + r * (uint64_t)rv.stride;
    uint64_t v = 0;
    if (p0 > -32) {
        const int q = p0 + 32, w = (q >> 5) - 1, s = (q & 31) * 2; // word w holds base p0 rounded down, shift s bits in
        const uint64_t hi = (w >= 0 && w < rv.stride) ? __ldg(src + w) : 0ULL;
        const uint64_t lo = (w + 1 >= 0 && w + 1 < rv.stride) ? __ldg(src + w + 1) : 0ULL;
        v = s ? ((hi << s) | (lo >> (64 - s))) : hi;
    }
    out[i] = v;
}

__global__ void k_rebase_rowinfo(uint64_t *rowinfo, uint64_t lo, uint64_t hi, uint64_t base)
{
    const uint64_t i = lo + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < hi) {
        const uint64_t ri = rowinfo[i];
        if (rowinfo_deg(ri)) rowinfo[i] = make_rowinfo(rowinfo_start(ri) + base, rowinfo_deg(ri));
    }
}

// the buffer that holds the row of read v: local, or (range-partitioned adjacency) the owning GPU's through NVLink
template <bool SHARDED>
__device__ __forceinline__ const uint64_t *rows_of(const ReduceParams &p, uint64_t v)
{
    if (!SHARDED) return p.rows;
    uint32_t r = 0;
    for (uint32_t i = 1; i < p.world; i++) r += (v >= p.bounds[i]) ? 1u : 0u;
    return p.peer_rows[r];
}

// ---------------------------------------------------------------------------------------------------------------
// transitive reduction, pass 1: Myers marking of every node on the full graph (OverlapGraph.cpp:687-723).
// Warp per node u; u's row lives in shared memory (neighbour id + orientation + state); for every neighbour v still
// INPLAY, in ascending offset order, the warp streams v's row (coalesced) and eliminates common neighbours whose
// orientations chain through v.  The result is the eliminated bit of u's own entries.
// ---------------------------------------------------------------------------------------------------------------
template <bool SHARDED>
__global__ void __launch_bounds__(kThreads, 8) k_reduce_mark(ReduceParams p) // 32 registers: full occupancy, the kernel waits on dependent row fetches
{
    extern __shared__ uint64_t smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    // per warp: ent u64[maxdeg] (the row; the entry value is also the (offset, id, orientation) sort key),
    //           hs  u32[hset]   (open-addressing set: neighbour id -> position in ent, sized 2x the row),
    //           st  u8[maxdeg]  (bit 0 visited, bit 1 eliminated; 0 = INPLAY and not yet visited)
    const int hcap = reduce_mark_hset(p.maxdeg);
    const size_t per_warp = reduce_mark_smem_per_warp(p.maxdeg);
    uint8_t *basep = reinterpret_cast<uint8_t *>(smem) + wib * per_warp;
    uint64_t *ent = reinterpret_cast<uint64_t *>(basep);
    uint32_t *hs = reinterpret_cast<uint32_t *>(basep + (size_t)p.maxdeg * 8);
    volatile uint8_t *st = basep + (size_t)p.maxdeg * 8 + (size_t)hcap * 4;
    unsigned long long n_rows = 0, n_ent = 0;
    uint64_t ub, ue;
    while (grab_chunk(p.work_counter, p.u_lo, p.u_hi, lane, &ub, &ue)) {
        for (uint64_t u = ub; u < ue; u++) {
            const uint64_t ri = p.rowinfo[u];
            const int deg = (int)rowinfo_deg(ri);
            if (deg == 0) continue;
            const uint64_t start = rowinfo_start(ri);
            int hsz = 64;
            while (hsz < 2 * deg) hsz <<= 1; // <= hcap
            const uint32_t hmask = (uint32_t)hsz - 1;
            for (int k = lane; k < hsz; k += 32) hs[k] = 0xFFFFFFFFu;
            __syncwarp();
            for (int k = lane; k < deg; k += 32) {
                const uint64_t e = __ldcg(p.rows + start + k) & ~kElimBit;
                ent[k] = e;
                st[k] = 0;
                uint32_t h = ((uint32_t)entry_nbr(e) * 0x9E3779B1u) >> 7 & hmask;
                while (atomicCAS(&hs[h], 0xFFFFFFFFu, (uint32_t)k) != 0xFFFFFFFFu) h = (h + 1) & hmask;
            }
            __syncwarp();
            for (;;) {
                // "traverse the list of edges according to their overlap offset" (OverlapGraph.cpp:693): rows are
                // stored unsorted, so pick the smallest (offset, id, orientation) entry that is still INPLAY and
                // unvisited -- a warp arg-min per visited neighbour (only a handful per node) instead of a sort
                uint64_t best = ~0ULL;
                int bk = -1;
                for (int k = lane; k < deg; k += 32) {
                    const uint64_t e = ent[k];
                    if (st[k] == 0 && e < best) { best = e; bk = k; }
                }
                {   // warp arg-min in two 32-bit hardware reductions (REDUX): offset first, then neighbour id -- one
                    // entry per neighbour, so (offset, id) is unique
                    const unsigned off = (best == ~0ULL) ? 0xFFFFFFFFu : (unsigned)entry_offset(best);
                    const unsigned moff = __reduce_min_sync(FULL, off);
                    if (moff == 0xFFFFFFFFu) break;
                    const unsigned id = (off == moff) ? (unsigned)entry_nbr(best) : 0xFFFFFFFFu;
                    const unsigned mid = __reduce_min_sync(FULL, id);
                    const int src = __ffs(__ballot_sync(FULL, off == moff && id == mid)) - 1;
                    best = __shfl_sync(FULL, best, src);
                    bk = __shfl_sync(FULL, bk, src);
                }
                if (best == ~0ULL) break;
                if (lane == 0) st[bk] = 1;
                const int t1 = entry_orient(best);
                const uint64_t vri = p.rowinfo[entry_nbr(best)];
                const int vd = (int)rowinfo_deg(vri);
                const uint64_t *vrow = rows_of<SHARDED>(p, entry_nbr(best)) + rowinfo_start(vri);
                n_rows += (lane == 0); n_ent += (lane == 0) ? (unsigned long long)vd : 0ULL;
                __syncwarp();
                for (int q = lane; q < vd; q += 32) {
                    const uint64_t e = (SHARDED && p.peer_load == 1) ? __ldg(vrow + q) : __ldcg(vrow + q); // other warps may be setting eliminated bits: ignored
                    if (!chain_ok(t1, entry_orient(e))) continue;
                    // is w also a neighbour of u?  (markedNodes->find(read3), OverlapGraph.cpp:701)
                    const uint32_t w = (uint32_t)entry_nbr(e);
                    uint32_t h = (w * 0x9E3779B1u) >> 7 & hmask;
                    for (;;) {
                        const uint32_t k = hs[h];
                        if (k == 0xFFFFFFFFu) break;
                        if ((uint32_t)entry_nbr(ent[k]) == w) { st[k] = st[k] | 2; break; } // ELIMINATED
                        h = (h + 1) & hmask;
                    }
                }
                __syncwarp();
            }
            for (int k = lane; k < deg; k += 32)
                if (st[k] & 2) p.rows[start + k] = ent[k] | kElimBit;
            __syncwarp();
        }
    }
    warp_stat_add(p.stats, ST_ROWS_FETCHED, n_rows);
    warp_stat_add(p.stats, ST_ENTRIES_FETCHED, n_ent);
}

// ---------------------------------------------------------------------------------------------------------------
// transitive reduction, pass 2: an edge dies when it was eliminated from either endpoint (edge and twin are flagged
// together in the reference, OverlapGraph.cpp:717-718).  For each surviving entry u->w the warp finds the twin in
// w's row; the lower id emits the canonical record (OverlapGraph.cpp:808).
// ---------------------------------------------------------------------------------------------------------------
template <bool SHARDED>
__global__ void __launch_bounds__(kThreads, 8) k_reduce_emit(ReduceParams p) // 32 registers: all 64 warps of an SM resident (latency bound)
{
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    disco_edge *out = reinterpret_cast<disco_edge *>(p.edges_out);
    disco_edge *out2 = reinterpret_cast<disco_edge *>(p.edges_out2);
    __shared__ disco_edge ebuf_all[kWarps][32]; // kept edges are staged per warp: one global atomic per 32 edges
    disco_edge *ebuf = ebuf_all[wib];
    int nbuf = 0;                               // warp-uniform
    unsigned long long n_rows = 0, n_ent = 0, n_multi = 0, n_one = 0, n_out = 0;
    uint64_t ub, ue;
    while (grab_chunk(p.work_counter, p.u_lo, p.u_hi, lane, &ub, &ue)) {
        for (uint64_t u = ub; u < ue; u++) {
            const uint64_t ri = p.rowinfo[u];
            const int deg = (int)rowinfo_deg(ri);
            if (deg == 0) continue;
            const uint64_t start = rowinfo_start(ri);
            if (u + 1 < ue) prefetch_row(p.rows, p.rowinfo[u + 1], lane);
            const int Lu = read_len(p.reads, u);
            for (int k0 = 0; k0 < deg; k0 += 32) {
                const int k = k0 + lane;
                const uint64_t e = (k < deg) ? p.rows[start + k] : kElimBit;
                unsigned alive = __ballot_sync(FULL, !(e & kElimBit));
                // the surviving entries' row infos are fetched together (one lane each) and their rows pulled into L2
                // before the twins are looked up one after the other
                uint64_t my_ri = 0;
                if (!(e & kElimBit)) my_ri = p.rowinfo[entry_nbr(e)];
                if (!SHARDED)
                    for (unsigned m = alive; m; m &= m - 1) prefetch_row(p.rows, __shfl_sync(FULL, my_ri, __ffs(m) - 1), lane);
                while (alive) {
                    const int src = __ffs(alive) - 1; alive &= alive - 1;
                    const uint64_t ee = __shfl_sync(FULL, e, src);
                    const uint64_t w = entry_nbr(ee);
                    const int orient = entry_orient(ee), offset = entry_offset(ee);
                    const uint64_t wri = __shfl_sync(FULL, my_ri, src);
                    const int wd = (int)rowinfo_deg(wri);
                    const uint64_t *wrow = rows_of<SHARDED>(p, w) + rowinfo_start(wri);
                    const int Lw = read_len(p.reads, w);
                    n_rows += (lane == 0); n_ent += (lane == 0) ? (unsigned long long)wd : 0ULL;
                    // twin of u->w as seen from w (OverlapGraph.cpp:617-619)
                    const int t_orient = twin_orient(orient), t_offset = Lw + offset - Lu;
                    int found = 0, dead = 0, same = 0;
                    for (int q0 = 0; q0 < wd; q0 += 32) {
                        const int q = q0 + lane;
                        uint64_t te = 0;
                        bool hit = false;
                        if (q < wd) { te = (SHARDED && p.peer_load == 1) ? __ldg(wrow + q) : __ldcg(wrow + q); hit = entry_nbr(te) == u; }
                        const unsigned hm = __ballot_sync(FULL, hit);
                        if (hm) {
                            const uint64_t t = __shfl_sync(FULL, te, __ffs(hm) - 1);
                            found = 1; dead = (t & kElimBit) != 0;
                            same = entry_orient(t) == t_orient && entry_offset(t) == t_offset;
                            break;
                        }
                    }
                    bool emit;
                    if (found) {
                        if (!same && u < w) n_multi += (lane == 0);
                        emit = !dead && u < w; // lower id's overlap is the canonical one
                    } else {
                        n_one += (lane == 0);
                        emit = true;           // the other endpoint cannot see this edge: emit it from here
                    }
                    if (emit) { // warp-uniform
                        if (lane == 0) {
                            disco_edge o;
                            if (u < w) { o.src = (uint32_t)u; o.dst = (uint32_t)w; o.offset = (uint32_t)offset; o.orient = (uint32_t)orient; }
                            else { o.src = (uint32_t)w; o.dst = (uint32_t)u; o.offset = (uint32_t)t_offset; o.orient = (uint32_t)t_orient; }
                            ebuf[nbuf] = o;
                            n_out++;
                        }
                        nbuf++;
                        if (nbuf == 32) {
                            __syncwarp();
                            unsigned long long pos = 0;
                            if (lane == 0) pos = atomicAdd(p.edges_cursor, 32ULL);
                            pos = __shfl_sync(FULL, pos, 0);
                            if (pos + lane < p.edges_cap) out[pos + lane] = ebuf[lane];
                            if (out2 && pos + lane < p.edges_cap2) out2[pos + lane] = ebuf[lane];
                            __syncwarp();
                            nbuf = 0;
                        }
                    }
                }
            }
        }
    }
    if (nbuf) {
        __syncwarp();
        unsigned long long pos = 0;
        if (lane == 0) pos = atomicAdd(p.edges_cursor, (unsigned long long)nbuf);
        pos = __shfl_sync(FULL, pos, 0);
        if (lane < nbuf && pos + lane < p.edges_cap) out[pos + lane] = ebuf[lane];
        if (out2 && lane < nbuf && pos + lane < p.edges_cap2) out2[pos + lane] = ebuf[lane];
    }
    warp_stat_add(p.stats, ST_EMIT_ROWS, n_rows);
    warp_stat_add(p.stats, ST_EMIT_ENTRIES, n_ent);
    warp_stat_add(p.stats, ST_MULTI_OVERLAP, n_multi);
    warp_stat_add(p.stats, ST_ONE_SIDED, n_one);
    warp_stat_add(p.stats, ST_EDGES_OUT, n_out);
}

// ---------------------------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------------------------
static int wp_of(int max_len) { return ((max_len + 31) >> 5) + 2; }

static std::atomic<unsigned long long> g_launches{0};
unsigned long long launches_total() { return g_launches.load(); }
void count_launches(unsigned long long k) { g_launches.fetch_add(k, std::memory_order_relaxed); }
#define DISCO_COUNT_LAUNCH() g_launches.fetch_add(1, std::memory_order_relaxed)

constexpr size_t kSmemBudget = 200 * 1024; // per block; leaves room for the driver's reservation out of 227 KB

// largest warp count (8, 4, 2, 1) whose shared memory fits the budget; 0 = even one warp does not fit
static int warps_that_fit(size_t per_warp_bytes)
{
    for (int w = kWarps; w >= 1; w >>= 1)
        if (per_warp_bytes * w <= kSmemBudget) return w;
    return 0;
}

template <typename Kern>
static cudaError_t persistent_grid(Kern kern, size_t smem, int num_sms, int *grid, int threads = kThreads)
{
    cudaError_t e = cudaSuccess;
    if (smem > 48 * 1024) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorInvalidConfiguration;
    *grid = num_sms * per_sm; // persistent: every CTA resident, work handed out by the atomic counter
    return cudaSuccess;
}

constexpr size_t kLaneSmemPerWarp = 12 * 1024; // lane-per-read kernels are used while 32 private array pairs fit in this

cudaError_t launch_table_insert(const ReadsView &r, const TableView &t, int K, const uint32_t *skip_bits,
                                int num_sms, cudaStream_t s, uint64_t r_lo, uint64_t r_hi, const unsigned int *gate)
{
    if (r_hi > r.n) r_hi = r.n; // (default: all reads)
    if (r_lo >= r_hi) return cudaSuccess;
    const uint64_t nr = r_hi - r_lo;
    {
        const size_t per_warp = (size_t)64 * lane_array_u32(r.max_len) * sizeof(uint32_t);
        if (per_warp <= kLaneSmemPerWarp) {
            const size_t smem = per_warp * kWarps;
            int grid = 0;
            cudaError_t e = persistent_grid(k_table_insert_lanes, smem, num_sms, &grid);
            if (e != cudaSuccess) return e;
            const uint64_t need = (nr + kWarps * 32 - 1) / (kWarps * 32);
            if ((uint64_t)grid > need) grid = (int)(need ? need : 1);
            k_table_insert_lanes<<<grid, kThreads, smem, s>>>(r, t, K, skip_bits, r_lo, r_hi, gate);
            DISCO_COUNT_LAUNCH();
            return cudaGetLastError();
        }
    }
    const size_t per_warp = 2 * (size_t)wp_of(r.max_len) * sizeof(uint64_t);
    const int warps = warps_that_fit(per_warp);
    if (!warps) return cudaErrorInvalidConfiguration;
    const size_t smem = per_warp * warps;
    int grid = 0;
    cudaError_t e = persistent_grid(k_table_insert, smem, num_sms, &grid, warps * 32);
    if (e != cudaSuccess) return e;
    const uint64_t need = (nr + warps - 1) / warps;
    if ((uint64_t)grid > need) grid = (int)(need ? need : 1);
    k_table_insert<<<grid, warps * 32, smem, s>>>(r, t, K, skip_bits, r_lo, r_hi, gate);
    DISCO_COUNT_LAUNCH();
    return cudaGetLastError();
}

static int pow2_at_least(int x) { int v = 1; while (v < x) v <<= 1; return v; }

// fills hcap / hset / rowcap for a launch: rowcap is the worst-case row (cap entries per position) and is also
// large enough to hold the fast path's scratch (id set of hset u32 + one int counter per position)
static void size_search(SearchParams &p, int mode)
{
    const int positions = p.reads.max_len - p.K;
    p.npos = positions;
    if (mode == MODE_EDGES) {
        const int worst = p.cap * positions;
        p.hcap = worst < kHitCap ? worst : kHitCap;
        p.hset = pow2_at_least(2 * p.hcap);
        const int scratch = (p.hset + positions + 2 + 1) / 2; // u64 units
        p.rowcap = worst > scratch ? worst : scratch;
    } else {
        p.hcap = kContainQueue; p.hset = 0; p.rowcap = 0;
    }
}

static size_t search_smem_per_warp(const SearchParams &p, int mode)
{
    return search_words_per_warp(wp_of(p.reads.max_len), p.npos, p.hcap, p.rowcap, mode) * sizeof(uint64_t);
}

bool search_edges_fits(int max_len, int K, int cap)
{
    SearchParams p{};
    p.reads.max_len = max_len; p.K = K; p.cap = cap;
    size_search(p, MODE_EDGES);
    return warps_that_fit(search_smem_per_warp(p, MODE_EDGES)) > 0;
}

template <int MODE>
static cudaError_t launch_search(const SearchParams &p_in, int num_sms, cudaStream_t s)
{
    SearchParams p = p_in;
    size_search(p, MODE);
    p.dbg = getenv("DISCO_DBG") ? atoi(getenv("DISCO_DBG")) : 0;
    const size_t per_warp = search_smem_per_warp(p, MODE);
    const int warps = warps_that_fit(per_warp);
    if (!warps) return cudaErrorInvalidConfiguration;
    const size_t smem = per_warp * warps;
    const int threads = warps * 32;
    int grid = 0;
    cudaError_t e;
#define DISCO_LAUNCH(NWV)                                                            \
    {                                                                                \
        e = persistent_grid(k_search<NWV, MODE>, smem, num_sms, &grid, threads);     \
        if (e != cudaSuccess) return e;                                              \
        k_search<NWV, MODE><<<grid, threads, smem, s>>>(p);                          \
        DISCO_COUNT_LAUNCH();                                                        \
        break;                                                                       \
    }
    const int words = (p.reads.max_len + 31) / 32; // registers hold an even number of words >= this
    switch (words <= 16 ? (words + 1) / 2 : 0) {
    case 1: DISCO_LAUNCH(2)
    case 2: DISCO_LAUNCH(4)
    case 3: DISCO_LAUNCH(6)
    case 4: DISCO_LAUNCH(8)
    case 5: case 6: DISCO_LAUNCH(12)
    case 7: case 8: DISCO_LAUNCH(16)
    default: DISCO_LAUNCH(0)
    }
#undef DISCO_LAUNCH
    return cudaGetLastError();
}

template <typename Kern>
static cudaError_t launch_warps(Kern kern, const SearchParams &p, size_t per_warp_bytes, int num_sms, cudaStream_t s)
{
    const int warps = warps_that_fit(per_warp_bytes);
    if (!warps) return cudaErrorInvalidConfiguration;
    int grid = 0;
    cudaError_t e = persistent_grid(kern, per_warp_bytes * warps, num_sms, &grid, warps * 32);
    if (e != cudaSuccess) return e;
    kern<<<grid, warps * 32, per_warp_bytes * warps, s>>>(p);
    DISCO_COUNT_LAUNCH();
    return cudaGetLastError();
}

bool table_bin_supported(int max_len)
{
    return (size_t)64 * lane_array_u32(max_len) * sizeof(uint32_t) <= kLaneSmemPerWarp;
}

cudaError_t launch_table_bin(const ReadsView &r, const TableView &t, int K, const BinView &b, int num_sms, cudaStream_t s,
                             uint64_t r_lo, uint64_t r_hi)
{
    if (r_hi > r.n) r_hi = r.n;
    if (r_lo >= r_hi) return cudaSuccess;
    const size_t lanes = ((size_t)kWarps * 64 * lane_array_u32(r.max_len) * sizeof(uint32_t) + 7) / 8 * 8;
    const size_t smem = lanes + (size_t)b.nbins * (sizeof(unsigned long long) + sizeof(uint32_t)) + 8;
    int grid = 0;
    cudaError_t e = persistent_grid(k_table_bin, smem, num_sms, &grid);
    if (e != cudaSuccess) return e;
    const uint64_t need = (r_hi - r_lo + kThreads - 1) / kThreads;
    if ((uint64_t)grid > need) grid = (int)need;
    k_table_bin<<<grid, kThreads, smem, s>>>(r, t, K, b, r_lo, r_hi);
    DISCO_COUNT_LAUNCH();
    return cudaGetLastError();
}

cudaError_t launch_table_fill(const TableView &t, const BinView &b, const uint32_t *skip_bits, int set_filter,
                              unsigned long long *work_counter, uint64_t n_unbinned, int num_sms, cudaStream_t s)
{
    int per_sm = 8;
    uint32_t chunk = 256;
    if (const char *e = getenv("DISCO_FILL_BLOCKS")) per_sm = std::max(1, std::min(8, atoi(e)));
    if (const char *e = getenv("DISCO_FILL_CHUNK")) chunk = (uint32_t)std::max(32, std::min(65536, atoi(e)));
    const uint32_t chunks_per_bin = (uint32_t)((b.cap + chunk - 1) / chunk);
    k_table_fill<<<num_sms * per_sm, kThreads, 0, s>>>(t, b, skip_bits, set_filter, work_counter, chunks_per_bin, chunk, n_unbinned);
    DISCO_COUNT_LAUNCH();
    return cudaGetLastError();
}

bool edges_flat_supported(int max_len, int stride, int K)
{
    if (getenv("DISCO_LEGACY_EDGES") || getenv("DISCO_FUSED")) return false; // the warp-per-read kernels, for A/B timing
    return stride <= 8 && max_len <= 32 * stride && (K + 31) / 32 <= 4;
}

cudaError_t launch_search_contained(const SearchParams &p, int num_sms, cudaStream_t s)
{
    const size_t per_warp = (size_t)64 * lane_array_u32(p.reads.max_len) * sizeof(uint32_t);
    const int words = (p.reads.max_len + 31) / 32;
    if (p.reads.uniform_len && per_warp <= kLaneSmemPerWarp && words <= 16) {
        const size_t smem = per_warp * kWarps;
        int grid = 0;
        cudaError_t e;
#define DISCO_LAUNCH_CU(NWV)                                                         \
    {                                                                                \
        e = persistent_grid(k_contain_uniform<NWV>, smem, num_sms, &grid);           \
        if (e != cudaSuccess) return e;                                              \
        k_contain_uniform<NWV><<<grid, kThreads, smem, s>>>(p);                      \
        DISCO_COUNT_LAUNCH();                                                        \
        break;                                                                       \
    }
        switch ((words + 1) / 2) {
        case 1: DISCO_LAUNCH_CU(2)
        case 2: DISCO_LAUNCH_CU(4)
        case 3: DISCO_LAUNCH_CU(6)
        case 4: DISCO_LAUNCH_CU(8)
        case 5: case 6: DISCO_LAUNCH_CU(12)
        default: DISCO_LAUNCH_CU(16)
        }
#undef DISCO_LAUNCH_CU
        return cudaGetLastError();
    }
    if (p.cands && edges_flat_supported(p.reads.max_len, p.reads.stride, p.K)) {
        // flat pass (edges_flat.cuh): probe<CONTAIN> lists the candidates of every 32-read batch, the verify kernel tests
        // them one per lane; batches it could not list go through the warp-per-read kernel
        const size_t pb = probe_flat_words_per_warp(p.reads.stride) * sizeof(uint64_t);
        const size_t vb = contain_flat_words_per_warp(p.reads.max_len) * sizeof(uint64_t);
        const bool sharded = p.table.world > 1;
        cudaError_t e;
#define DISCO_LAUNCH_PC(KWV)                                                                                      \
    e = sharded ? launch_warps(k_probe_flat<KWV, true, true>, p, pb, num_sms, s) : launch_warps(k_probe_flat<KWV, false, true>, p, pb, num_sms, s); \
    break;
        switch ((p.K + 31) / 32) {
        case 1: DISCO_LAUNCH_PC(1)
        case 2: DISCO_LAUNCH_PC(2)
        case 3: DISCO_LAUNCH_PC(3)
        default: DISCO_LAUNCH_PC(4)
        }
#undef DISCO_LAUNCH_PC
        if (e != cudaSuccess) return e;
        switch ((words + 1) / 2) {
        case 1: e = launch_warps(k_contain_verify_flat<2>, p, vb, num_sms, s); break;
        case 2: e = launch_warps(k_contain_verify_flat<4>, p, vb, num_sms, s); break;
        case 3: e = launch_warps(k_contain_verify_flat<6>, p, vb, num_sms, s); break;
        default: e = launch_warps(k_contain_verify_flat<8>, p, vb, num_sms, s); break;
        }
        if (e != cudaSuccess) return e;
        SearchParams q = p;
        q.only_flagged = 1;
        q.work_counter = p.work_counter + 2;
        return launch_search<MODE_CONTAIN>(q, num_sms, s);
    }
    return launch_search<MODE_CONTAIN>(p, num_sms, s);
}


uint64_t edges_flat_slack(int num_sms) { return (uint64_t)num_sms * 5 * kWarps * kFlatSlice; }

// probe (one lane per read, then one lane per probe) -> verify (one lane per candidate) -> exact (flagged reads)
static cudaError_t launch_edges_flat(SearchParams &p, int num_sms, cudaStream_t s, cudaEvent_t ev_probe_done, cudaEvent_t ev_verify_done)
{
    const int WP = wp_of(p.reads.max_len);
    const size_t pb = probe_flat_words_per_warp(p.reads.stride) * sizeof(uint64_t);
    const size_t vb = verify_flat_words_per_warp(p.reads.max_len) * sizeof(uint64_t);
    const size_t xb = exact_words_per_warp(WP, p.rowcap) * sizeof(uint64_t);
    const bool sharded = p.table.world > 1;
    const bool sect = p.reads.tails != nullptr && p.reads.uniform_len > 128 && p.reads.stride == 8; // two sectors per row
    cudaError_t e;
#define DISCO_LAUNCH_P(KWV)                                                                                       \
    e = sharded ? launch_warps(k_probe_flat<KWV, true, false>, p, pb, num_sms, s) : launch_warps(k_probe_flat<KWV, false, false>, p, pb, num_sms, s); \
    break;
    switch ((p.K + 31) / 32) {
    case 1: DISCO_LAUNCH_P(1)
    case 2: DISCO_LAUNCH_P(2)
    case 3: DISCO_LAUNCH_P(3)
    default: DISCO_LAUNCH_P(4)
    }
#undef DISCO_LAUNCH_P
    if (e != cudaSuccess) return e;
    if (ev_probe_done) cudaEventRecord(ev_probe_done, s);
#define DISCO_LAUNCH_V(NWV)                                                          \
    {                                                                                \
        e = sect ? launch_warps(k_verify_flat<NWV, true>, p, vb, num_sms, s)         \
                 : launch_warps(k_verify_flat<NWV, false>, p, vb, num_sms, s);       \
        if (e != cudaSuccess) return e;                                              \
        if (ev_verify_done) cudaEventRecord(ev_verify_done, s);                      \
        e = launch_warps(k_edges_exact<NWV>, p, xb, num_sms, s);                     \
        break;                                                                       \
    }
    switch (((p.reads.max_len + 31) / 32 + 1) / 2) {
    case 1: DISCO_LAUNCH_V(2)
    case 2: DISCO_LAUNCH_V(4)
    case 3: DISCO_LAUNCH_V(6)
    default: DISCO_LAUNCH_V(8)
    }
#undef DISCO_LAUNCH_V
    return e;
}

cudaError_t launch_search_edges(const SearchParams &p_in, int num_sms, cudaStream_t s, cudaEvent_t ev_probe_done, cudaEvent_t ev_verify_done)
{
    if (getenv("DISCO_FUSED")) return launch_search<MODE_EDGES>(p_in, num_sms, s); // the single-kernel variant, for A/B timing
    SearchParams p = p_in;
    size_search(p, MODE_EDGES);
    if (p.cands && edges_flat_supported(p.reads.max_len, p.reads.stride, p.K)) return launch_edges_flat(p, num_sms, s, ev_probe_done, ev_verify_done);
    const int WP = wp_of(p.reads.max_len);
    const size_t pb = probe_words_per_warp(WP, p.npos, p.hcap) * sizeof(uint64_t);
    cudaError_t e = p.table.world > 1 ? launch_warps(k_edges_probe<true>, p, pb, num_sms, s) : launch_warps(k_edges_probe<false>, p, pb, num_sms, s);
    if (e != cudaSuccess) return e;
    if (ev_probe_done) cudaEventRecord(ev_probe_done, s);
    // short reads: two 16-lane groups per warp in the verify kernel (DISCO_VERIFY_GW=32 selects one read per warp)
    const int gw = (getenv("DISCO_VERIFY_GW") ? atoi(getenv("DISCO_VERIFY_GW")) : 16) == 16 && (p.reads.max_len + 31) / 32 <= 8 ? 16 : 32;
    const size_t vb = verify_words_per_group(WP, p.npos, p.hcap, p.hset, gw) * sizeof(uint64_t) * (32 / gw);
    const size_t xb = exact_words_per_warp(WP, p.rowcap) * sizeof(uint64_t);
#define DISCO_LAUNCH_E(NWV)                                                          \
    {                                                                                \
        e = gw == 16 ? launch_warps(k_edges_verify<NWV, (NWV > 0 && NWV <= 8) ? 16 : 32>, p, vb, num_sms, s)   \
                     : launch_warps(k_edges_verify<NWV, 32>, p, vb, num_sms, s);     \
        if (e != cudaSuccess) return e;                                              \
        if (ev_verify_done) cudaEventRecord(ev_verify_done, s);                      \
        e = launch_warps(k_edges_exact<NWV>, p, xb, num_sms, s);                     \
        break;                                                                       \
    }
    const int words = (p.reads.max_len + 31) / 32;
    switch (words <= 16 ? (words + 1) / 2 : 0) {
    case 1: DISCO_LAUNCH_E(2)
    case 2: DISCO_LAUNCH_E(4)
    case 3: DISCO_LAUNCH_E(6)
    case 4: DISCO_LAUNCH_E(8)
    case 5: case 6: DISCO_LAUNCH_E(12)
    case 7: case 8: DISCO_LAUNCH_E(16)
    default: DISCO_LAUNCH_E(0)
    }
#undef DISCO_LAUNCH_E
    return e;
}

cudaError_t launch_contained_finish(const unsigned long long *best, uint64_t n, uint32_t *bits, unsigned long long *count, cudaStream_t s)
{
    if (n == 0) return cudaSuccess;
    const uint64_t blocks = (n + 255) / 256;
    k_contained_finish<<<(unsigned)blocks, 256, 0, s>>>(best, n, bits, count);
    DISCO_COUNT_LAUNCH();
    return cudaGetLastError();
}

cudaError_t launch_contained_rows(const unsigned long long *best, const ReadsView &r, int K, void *rows_out,
                                  unsigned long long *cursor, cudaStream_t s)
{
    if (r.n == 0) return cudaSuccess;
    const uint64_t blocks = (r.n + 255) / 256;
    k_contained_rows<<<(unsigned)blocks, 256, 0, s>>>(best, r, K, reinterpret_cast<disco_crow *>(rows_out), cursor);
    DISCO_COUNT_LAUNCH();
    return cudaGetLastError();
}

cudaError_t launch_crows_in_range(const void *in, uint64_t n, uint64_t lo, uint64_t hi, void *out, unsigned long long *cursor, cudaStream_t s)
{
    if (n == 0) return cudaSuccess;
    k_crows_in_range<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(reinterpret_cast<const disco_crow *>(in), n, (uint32_t)lo, (uint32_t)std::min<uint64_t>(hi, 0xFFFFFFFFULL),
                                                                reinterpret_cast<disco_crow *>(out), cursor);
    DISCO_COUNT_LAUNCH();
    return cudaGetLastError();
}

cudaError_t launch_restride(const uint64_t *src, int src_stride, int src_words, uint64_t *dst, int dst_stride, uint64_t n, cudaStream_t s)
{
    if (n == 0) return cudaSuccess;
    const uint64_t total = n * (uint64_t)dst_stride;
    k_restride<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(src, src_stride, src_words, dst, dst_stride, n);
    DISCO_COUNT_LAUNCH();
    return cudaGetLastError();
}

// in-place MIN over the ranks' containment keys (single-process multi-GPU: the peers' arrays are read through NVLink).
// Every rank runs this on its own array at the same time; values only ever decrease towards the global minimum, so a
// peer's half-reduced value is as good as its original one.
__global__ void k_min_keys(unsigned long long *mine, const uint64_t *const *peers, uint32_t world, uint32_t rank, uint64_t n)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long v = mine[i];
    for (uint32_t r = 0; r < world; r++)
        if (r != rank) { const unsigned long long o = __ldcg(reinterpret_cast<const unsigned long long *>(peers[r]) + i); v = o < v ? o : v; }
    mine[i] = v;
}

// Containment keys that are set, as (read, key) pairs -- what a rank contributes to the exchange of the keys: a few
// percent of n instead of the whole array.  Warp-aggregated append; order is irrelevant (the consumer takes minima).
__global__ void k_compact_keys(const unsigned long long *best, uint64_t n, unsigned long long *pairs, uint64_t cap, unsigned long long *count)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long k = i < n ? best[i] : ~0ULL;
    const bool set = k != ~0ULL;
    const unsigned m = __ballot_sync(FULL, set);
    const int lane = threadIdx.x & 31;
    const unsigned long long base = block_reserve(m, count);
    if (set) {
        const unsigned long long at = base + __popc(m & ((1u << lane) - 1));
        if (at < cap) { pairs[2 * at] = i; pairs[2 * at + 1] = k; }
    }
}

// best[read] = min(best[read], key) for every pair (pairs with read >= n are padding)
__global__ void k_apply_keys(unsigned long long *best, uint64_t n, const unsigned long long *pairs, uint64_t npairs)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npairs) return;
    const unsigned long long r = pairs[2 * i], k = pairs[2 * i + 1];
    if (r < n && k != ~0ULL) atomicMin(best + r, k);
}

cudaError_t launch_compact_keys(const unsigned long long *best, uint64_t n, unsigned long long *pairs, uint64_t cap, unsigned long long *count, cudaStream_t s)
{
    if (n == 0) return cudaSuccess;
    k_compact_keys<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(best, n, pairs, cap, count);
    DISCO_COUNT_LAUNCH();
    return cudaGetLastError();
}

cudaError_t launch_apply_keys(unsigned long long *best, uint64_t n, const unsigned long long *pairs, uint64_t npairs, cudaStream_t s)
{
    if (npairs == 0) return cudaSuccess;
    k_apply_keys<<<(unsigned)((npairs + 255) / 256), 256, 0, s>>>(best, n, pairs, npairs);
    DISCO_COUNT_LAUNCH();
    return cudaGetLastError();
}

cudaError_t launch_min_keys(unsigned long long *mine, const uint64_t *const *peers, uint32_t world, uint32_t rank, uint64_t n, cudaStream_t s)
{
    if (n == 0) return cudaSuccess;
    k_min_keys<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(mine, peers, world, rank, n);
    DISCO_COUNT_LAUNCH();
    return cudaGetLastError();
}

cudaError_t launch_make_tails(const ReadsView &r, uint64_t *out, cudaStream_t s)
{
    if (r.n == 0) return cudaSuccess;
    k_make_tails<<<(unsigned)((r.n * 4 + 255) / 256), 256, 0, s>>>(r, out);
    DISCO_COUNT_LAUNCH();
    return cudaGetLastError();
}

cudaError_t launch_revcomp_rows(const ReadsView &r, uint64_t *out, cudaStream_t s)
{
    if (r.n == 0) return cudaSuccess;
    k_revcomp_rows<<<(unsigned)((r.n + 255) / 256), 256, 0, s>>>(r, out);
    DISCO_COUNT_LAUNCH();
    return cudaGetLastError();
}

cudaError_t launch_rebase_rowinfo(uint64_t *rowinfo, uint64_t u_lo, uint64_t u_hi, uint64_t base, cudaStream_t s)
{
    if (u_hi <= u_lo) return cudaSuccess;
    const uint64_t blocks = (u_hi - u_lo + 255) / 256;
    k_rebase_rowinfo<<<(unsigned)blocks, 256, 0, s>>>(rowinfo, u_lo, u_hi, base);
    DISCO_COUNT_LAUNCH();
    return cudaGetLastError();
}

cudaError_t launch_reduce_mark(const ReduceParams &p, int num_sms, cudaStream_t s)
{
    const size_t per_warp = reduce_mark_smem_per_warp(p.maxdeg);
    const int warps = warps_that_fit(per_warp);
    if (!warps) return cudaErrorInvalidConfiguration;
    const size_t smem = per_warp * warps;
    int grid = 0;
    auto kern = p.world > 1 ? k_reduce_mark<true> : k_reduce_mark<false>;
    cudaError_t e = persistent_grid(kern, smem, num_sms, &grid, warps * 32);
    if (e != cudaSuccess) return e;
    kern<<<grid, warps * 32, smem, s>>>(p);
    DISCO_COUNT_LAUNCH();
    return cudaGetLastError();
}

cudaError_t launch_reduce_emit(const ReduceParams &p, int num_sms, cudaStream_t s)
{
    int grid = 0;
    auto kern = p.world > 1 ? k_reduce_emit<true> : k_reduce_emit<false>;
    cudaError_t e = persistent_grid(kern, 0, num_sms, &grid);
    if (e != cudaSuccess) return e;
    kern<<<grid, kThreads, 0, s>>>(p);
    DISCO_COUNT_LAUNCH();
    return cudaGetLastError();
}

} // namespace disco
